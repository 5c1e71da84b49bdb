// fries_launch -n N <driver> [driver arguments]: starts N copies of a driver (frisys_mol, frifull_mol), one per GPU of this
// box -- the role of `mpirun -n N` for the reference's drivers (FRIES_bin/frisys_mol.cpp:62-63).  Every copy gets
// FRIES_NRANKS / FRIES_RANK (its GPU) / FRIES_RDV (a private directory through which the copies exchange CUDA IPC handles
// and the seed, host/fries_host.hpp: Ranks); everything per-iteration then happens inside kernels over NVLink.  Plain POSIX:
// no MPI, no Python.  Exit code: the first non-zero exit code of a rank (the others are then terminated), else 0.
#include <cerrno>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

int main(int argc, char *argv[]) {
    int n = 0, first = 1;
    if (argc >= 4 && std::strcmp(argv[1], "-n") == 0) {
        n = std::atoi(argv[2]);
        first = 3;
    }
    if (n < 1 || n > 8 || first >= argc) {
        std::fprintf(stderr, "usage: fries_launch -n <1..8> <driver> [arguments]\n");
        return 2;
    }
    char tmpl[] = "/tmp/fries_rdv_XXXXXX";
    const char *rdv = mkdtemp(tmpl);
    if (!rdv) {
        std::perror("fries_launch: mkdtemp");
        return 2;
    }
    std::vector<pid_t> pids;
    for (int r = 0; r < n; r++) {
        pid_t pid = fork();
        if (pid < 0) {
            std::perror("fries_launch: fork");
            for (pid_t p : pids) kill(p, SIGTERM);
            return 2;
        }
        if (pid == 0) {
            setenv("FRIES_NRANKS", std::to_string(n).c_str(), 1);
            setenv("FRIES_RANK", std::to_string(r).c_str(), 1);
            setenv("FRIES_RDV", rdv, 1);
            execvp(argv[first], argv + first);
            std::fprintf(stderr, "fries_launch: cannot start %s: %s\n", argv[first], std::strerror(errno));
            _exit(127);
        }
        pids.push_back(pid);
    }
    int rc = 0, left = n;
    while (left > 0) {
        int status = 0;
        pid_t p = wait(&status);
        if (p < 0) break;
        left--;
        int code = WIFEXITED(status) ? WEXITSTATUS(status) : 128 + (WIFSIGNALED(status) ? WTERMSIG(status) : 0);
        if (code != 0 && rc == 0) {
            rc = code;
            for (pid_t q : pids)
                if (q != p) kill(q, SIGTERM);  // a rank that lost its peers would wait for them forever
        }
    }
    std::string cmd = std::string("rm -rf '") + rdv + "'";
    if (std::system(cmd.c_str()) != 0) std::fprintf(stderr, "fries_launch: could not remove %s\n", rdv);
    return rc;
}
