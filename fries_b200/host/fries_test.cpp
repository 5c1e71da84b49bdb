// the reference's link smoke test (examples/fries_test.cpp): an external program links the host layer and calls
// round_binomially
#include "fries_host.hpp"
int main() {
    std::mt19937 mt(0);
    int r = fries::round_binomially(2.5, 10, mt);
    std::cout << "round_binomially(2.5, 10) = " << r << ", library version " << fries_version() << std::endl;
    return (r >= 20 && r <= 30) ? 0 : 1;
}
