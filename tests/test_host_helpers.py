"""Host-side helpers of libevk.so that need no GPU: the staging copy of the host pipeline (evk_host_copy: persistent
worker pool, non-temporal stores) and the content hash of an event set (evk_host_hash64[_multi]) that runs on the same
pool.  The reference has no counterpart -- its callers hand numpy arrays to torch (lib/data_loaders/base_dataset.py:446-453)."""
import ctypes
import threading

import numpy as np

from event_utils_b200 import _lib

VP = ctypes.c_void_p


def _copy(L, dsts, srcs, nbytes):
    k = len(dsts)
    L.evk_host_copy((VP * k)(*[a.ctypes.data for a in dsts]), (VP * k)(*[a.ctypes.data for a in srcs]), k, nbytes)


def test_host_copy_every_size_and_alignment():
    L = _lib.load()
    rng = np.random.default_rng(0)
    for n in (0, 1, 15, 16, 255, 256, 257, 4097, (1 << 20) - 1, (1 << 20) + 3, 5 * (1 << 20) + 77):
        for so, do in ((0, 0), (1, 0), (0, 3), (7, 16), (5, 9)):
            src = [rng.integers(0, 256, n + 64, dtype=np.uint8) for _ in range(4)]
            dst = [np.full(n + 64, 0xAB, np.uint8) for _ in range(4)]
            _copy(L, [d[do:] for d in dst], [s[so:] for s in src], n)
            for d, s in zip(dst, src):
                assert np.array_equal(d[do:do + n], s[so:so + n]), (n, so, do)
                assert (d[:do] == 0xAB).all() and (d[do + n:] == 0xAB).all(), (n, so, do)     # nothing outside the range


def test_host_pool_serves_concurrent_callers():
    """Copies and hashes from several Python threads at once (ctypes releases the GIL): runs on the pool are serialised,
    every caller gets its own result."""
    L = _lib.load()
    rng = np.random.default_rng(1)
    n = 6 * (1 << 20) + 5
    srcs = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(6)]
    want = [L.evk_host_hash64(s.ctypes.data, n, 7) for s in srcs]
    errors = []

    def worker(i):
        try:
            for _ in range(5):
                dst = np.zeros(n, np.uint8)
                _copy(L, [dst], [srcs[i]], n)
                assert np.array_equal(dst, srcs[i])
                assert L.evk_host_hash64(dst.ctypes.data, n, 7) == want[i]
        except Exception as e:          # noqa: BLE001 -- reported below
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(set(want)) == 6


def test_hash_does_not_depend_on_the_thread_count():
    """The same buffers hashed in a process whose pool has one thread (EVK_HOST_THREADS=1) give the same values."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = ("import sys; sys.path.insert(0, %r)\n"
              "import numpy as np\nfrom event_utils_b200 import _lib\n"
              "rng = np.random.default_rng(3)\n"
              "arrs = [rng.random(n) for n in (10, 300_000, 1_500_001)]\n"
              "print(*_lib.host_hashes(arrs))\n") % root
    outs = []
    for env in ({}, {"EVK_HOST_THREADS": "1"}, {"EVK_HOST_THREADS": "5"}):
        r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] == outs[2] and len(outs[0].split()) == 3


def test_forked_child_works_without_the_pool_threads():
    """os.fork() copies the pool object but not its threads (a DataLoader worker): the child must not wait for them."""
    import os
    L = _lib.load()
    rng = np.random.default_rng(4)
    n = 8 * (1 << 20) + 1
    src = rng.integers(0, 256, n, dtype=np.uint8)
    want = L.evk_host_hash64(src.ctypes.data, n, 1)          # the pool exists now
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        ok = b"0"
        try:
            dst = np.zeros(n, np.uint8)
            _copy(L, [dst], [src], n)
            if np.array_equal(dst, src) and L.evk_host_hash64(dst.ctypes.data, n, 1) == want:
                ok = b"1"
        finally:
            os.write(w, ok)
            os._exit(0)
    os.close(w)
    import select
    ready, _, _ = select.select([r], [], [], 120.0)
    if not ready:
        os.kill(pid, 9)
    os.waitpid(pid, 0)
    assert ready and os.read(r, 1) == b"1"


def test_host_pool_is_race_free_under_thread_sanitizer(tmp_path):
    """The HostPool class text, cut out of csrc/evk_host.cu, compiled with g++ -fsanitize=thread and driven by four
    concurrent callers with changing widths: every job runs exactly once and ThreadSanitizer reports nothing.
    Skipped where the toolchain has no TSan runtime."""
    import os
    import shutil
    import subprocess
    import pytest
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "event_utils_b200", "csrc", "evk_host.cu")).read()
    a = src.index("namespace {\nclass HostPool {")
    b = src.index("HostPool *HostPool::self_ = nullptr;")
    headers = "".join("#include <%s>\n" % h for h in ("pthread.h", "sched.h", "stdlib.h", "string.h", "stdint.h", "stdio.h", "atomic",
                                                      "condition_variable", "functional", "mutex", "thread", "vector"))
    harness = r'''
int main() {
    std::vector<std::thread> callers;
    std::atomic<long> total{0};
    for (int c = 0; c < 4; ++c) callers.emplace_back([&, c] {
        for (int rep = 0; rep < 300; ++rep) {
            std::vector<int> hit(257 + rep % 7, 0);
            const int width = 1 + (rep * 7 + c) % 12;
            HostPool::get().run(hit.size(), width, [&](size_t j) { hit[j] += 1; });
            for (size_t j = 0; j < hit.size(); ++j) if (hit[j] != 1) { printf("BAD %zu %d\n", j, hit[j]); exit(1); }
            total += (long)hit.size();
        }
    });
    for (auto &t : callers) t.join();
    printf("ok %ld\n", total.load());
    return 0;
}
'''
    cpp = tmp_path / "pool.cpp"
    cpp.write_text(headers + src[a:b] + "HostPool *HostPool::self_ = nullptr;\n}  // namespace\n" + harness)
    exe = tmp_path / "pool_tsan"
    cc = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", str(cpp), "-o", str(exe), "-lpthread"],
                        capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime here: " + cc.stderr[-300:])
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300, env=dict(os.environ, EVK_HOST_THREADS="8"))
    assert run.returncode == 0 and "ok " in run.stdout, run.stdout[-1000:] + run.stderr[-3000:]
    assert "ThreadSanitizer" not in run.stderr, run.stderr[-3000:]
