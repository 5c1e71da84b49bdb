#pragma once
#include "fries_global.hpp"
