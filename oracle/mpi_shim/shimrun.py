#!/usr/bin/env python3
"""Start N ranks of a program compiled against oracle/mpi_shim/mpi.h (TEST INFRASTRUCTURE ONLY).

    python oracle/mpi_shim/shimrun.py -n 8 [--slot-mb 64] [--timeout 600] -- oracle/_ref/frisys_mol --fcidump_path ...

The ranks share one memory-mapped file (under /dev/shm when it exists) that holds a barrier and one slot per rank; see
the header for the collectives.  Rank 0's stdout / stderr are passed through, the others' are discarded unless
--all-output is given.  If a rank dies or the timeout expires every rank is killed (a dead rank would leave the others
spinning in a barrier) and the exit code is non-zero."""
import argparse
import os
import signal
import struct
import subprocess
import sys
import tempfile
import time

HDR = 4096


def run(n, cmd, slot_bytes=64 << 20, timeout=None, env=None, all_output=False, stdout=None, stderr=None, stamp=None,
        grace=3.0, cwd=None):
    """returns (exit code of rank 0 or the first failing rank, wall seconds).  stamp: a compiled regex with one group;
    the ranks' stdout then goes through one pseudo-terminal (so that it is line buffered) and the arrival time of every
    matching line is appended to the list stamp_out as (group(1), perf_counter) -- pass it as run.stamps afterwards."""
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    fd, path = tempfile.mkstemp(prefix="fries_shim_", dir=base)
    procs = []
    run.stamps = []
    master = slave = None
    pending = b""
    if stamp is not None:
        import pty
        master, slave = pty.openpty()
    t0 = time.perf_counter()
    try:
        os.ftruncate(fd, HDR + n * slot_bytes)  # sparse: only touched pages are allocated
        os.pwrite(fd, struct.pack("iiiiQ", 0, 0, n, 0, slot_bytes), 0)
        os.close(fd)
        for r in range(n):
            e = dict(os.environ if env is None else env, FRIES_SHIM_SHM=path, FRIES_SHIM_RANK=str(r), FRIES_SHIM_SIZE=str(n),
                     FRIES_SHIM_SLOT=str(slot_bytes))
            quiet = r != 0 and not all_output
            # the drivers print their per-iteration line on the rank that owns the Hartree-Fock determinant, not on rank 0
            out = slave if slave is not None else (subprocess.DEVNULL if quiet else stdout)
            procs.append(subprocess.Popen(cmd, env=e, stdout=out, stderr=subprocess.DEVNULL if quiet else stderr,
                                          start_new_session=True, cwd=cwd))
        if slave is not None:
            os.close(slave)
            slave = None

        def drain(wait):
            nonlocal pending
            import select
            while True:
                ready, _, _ = select.select([master], [], [], wait)
                if not ready:
                    return
                try:
                    chunk = os.read(master, 65536)
                except OSError:
                    return
                if not chunk:
                    return
                now = time.perf_counter()
                pending += chunk
                *lines, pending = pending.split(b"\n")
                for ln in lines:
                    m = stamp.match(ln.decode(errors="replace").strip())
                    if m:
                        run.stamps.append((m.group(1), now))
                wait = 0

        rc = None
        failed_at = None
        while rc is None:
            if master is not None:
                drain(0.005)
            codes = [p.poll() for p in procs]
            if all(c is not None for c in codes):
                rc = next((c for c in codes if c), 0)
            elif any(c not in (None, 0) for c in codes):
                # a rank failed: the others may be about to finish too (e.g. a test binary that fails on every rank);
                # give them a moment before they are killed -- if they wait in a barrier they will never finish
                if failed_at is None:
                    failed_at = time.perf_counter()
                if time.perf_counter() - failed_at > grace:
                    rc = next(c for c in codes if c not in (None, 0))
                elif master is None:
                    time.sleep(0.01)
            elif timeout is not None and time.perf_counter() - t0 > timeout:
                rc = 124
            elif master is None:
                time.sleep(0.01)
        if master is not None:
            drain(0.05)
        return rc, time.perf_counter() - t0
    finally:
        for fdx in (master, slave):
            if fdx is not None:
                try:
                    os.close(fdx)
                except OSError:
                    pass
        for p in procs:
            if p.poll() is None:
                try:
                    os.killpg(p.pid, signal.SIGKILL)  # the rank's own session: exactly the process we started
                except ProcessLookupError:
                    pass
        for p in procs:
            try:
                p.wait(timeout=5)
            except Exception:
                pass
        try:
            os.unlink(path)
        except OSError:
            pass


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-n", type=int, required=True)
    ap.add_argument("--slot-mb", type=int, default=64)
    ap.add_argument("--timeout", type=float, default=None)
    ap.add_argument("--all-output", action="store_true")
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = a.cmd[1:] if a.cmd and a.cmd[0] == "--" else a.cmd
    rc, sec = run(a.n, cmd, a.slot_mb << 20, a.timeout, all_output=a.all_output)
    sys.exit(rc)


if __name__ == "__main__":
    main()
