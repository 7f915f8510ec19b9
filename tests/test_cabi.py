"""CPU-side boundary tests: the C-ABI library loads, exports every symbol include/tn_c_api.h
declares, and fails loudly (no CPU fallback) when no GPU is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tn_c_api.h")).read()
    return sorted(set(re.findall(r"\b(tn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import tnb200
    from tnb200 import _lib
    lib = tnb200.load()
    syms = _declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"missing export {s}"
    # every prototype the Python mirror binds is declared in the header
    for s in _lib.PROTOTYPES:
        assert s in syms
    assert lib.tn_version() >= 100


def test_no_cpu_fallback():
    import torch
    import tnb200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tnb200.TNError):
        tnb200.Context(0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensornetworks.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_null_handles_are_errors_not_crashes():
    """Every entry point that takes a handle rejects NULL with TN_ERR_INVALID and a message (no GPU needed: the check comes first)."""
    import ctypes as C
    import tnb200
    from tnb200 import _lib
    lib = tnb200.load()
    v = _lib.tn_cplx()
    k = C.c_int64()
    tr = _lib.tn_trunc_t(0.0, 0, 1)
    calls = [lambda: lib.tn_mps_norm(None, C.byref(v)), lambda: lib.tn_mps_normalize(None), lambda: lib.tn_mps_movecenter(None, 1, tr),
             lambda: lib.tn_env_movecenter(None, 1), lambda: lib.tn_env_calculate(None, C.byref(v)), lambda: lib.tn_envsum_movecenter(None, 1),
             lambda: lib.tn_apply_gates(None, None, tr), lambda: lib.tn_mps_maxbonddim(None, C.byref(k)), lambda: lib.tn_sync(None),
             lambda: lib.tn_mpo_compress(None, tr), lambda: lib.tn_heff_sharded_apply(None, None, None),
             lambda: lib.tn_svd_trunc_split(None, None, 4, 4, tr, 1, None, None, None, C.byref(k), None, 1, None)]
    for f in calls:
        assert f() == -1
        assert b"null" in lib.tn_last_error()
    # the free functions accept NULL like free()
    assert lib.tn_mps_free(None) == 0 and lib.tn_env_free(None) == 0 and lib.tn_gates_free(None) == 0 and lib.tn_envsum_free(None) == 0
    assert lib.tn_heff_sharded_free(None) == 0
    # the sharded matvec validates its device list before touching CUDA
    h = C.c_void_p()
    assert lib.tn_heff_sharded_create(0, None, 4, 4, 2, 3, 3, 3, None, None, None, None, _lib.tn_cplx(1.0, 0.0), C.byref(h)) != 0


def test_philox_known_answer_vectors_and_uniform_statistics():
    """The counter-based generator of the QJMC throughput runs (csrc/tn_qjmc.cu) is Philox4x32-10: the three Random123
    known-answer vectors, and the [0, 1) uniforms drawn for (seed, trajectory, step, slot): 53-bit resolution, mean / variance / a
    64-bin chi-square within sampling error, no correlation between neighbouring trajectories, steps or slots.  Host functions: no GPU."""
    import ctypes as C
    import numpy as np
    import tnb200
    lib = tnb200.load()
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        assert lib.tn_philox4x32_10(c, k, o) == 0
        assert list(o) == want

    def u(seed, traj, step, slot):
        out = C.c_double()
        assert lib.tn_qjmc_uniform(seed, traj, step, slot, C.byref(out)) == 0
        return out.value
    x = np.array([[[u(7, t, s, sl) for sl in range(3)] for s in range(40)] for t in range(200)])      # 24 000 uniforms
    assert x.min() >= 0.0 and x.max() < 1.0
    n = x.size
    assert abs(x.mean() - 0.5) < 5 * np.sqrt(1 / 12 / n)
    assert abs(x.var() - 1 / 12) < 5 * np.sqrt(1 / 180 / n)
    hist = np.bincount((x.reshape(-1) * 64).astype(int), minlength=64)
    chi2 = float(np.sum((hist - n / 64) ** 2 / (n / 64)))
    assert chi2 < 63 + 5 * np.sqrt(2 * 63), chi2
    for a, b in ((x[:-1], x[1:]), (x[:, :-1], x[:, 1:]), (x[:, :, :-1], x[:, :, 1:])):
        r = np.corrcoef(a.reshape(-1), b.reshape(-1))[0, 1]
        assert abs(r) < 5 / np.sqrt(a.size), r
    assert u(7, 3, 5, 1) == u(7, 3, 5, 1) and u(7, 3, 5, 1) != u(8, 3, 5, 1)       # a pure function of its arguments, keyed by the seed
    assert len({u(1, t, 0, 0) for t in range(1000)}) == 1000
    assert any((u(1, t, 0, 0) * 2 ** 53) % 2 == 1 for t in range(64))             # the lowest of the 53 bits is in use


def test_split_pair_schedule_visits_every_pair_once_with_disjoint_concurrent_tasks():
    """The Jacobi sweeps of large factorisations (>= 32 column blocks) run a pair schedule whose phases fall apart into 2 or 4 tasks on
    separate streams (csrc/tn_svd.cu "Split schedule").  For every block count the library may see: every unordered pair of blocks exactly
    once per sweep, the pairs of a step disjoint, the tasks of a phase on disjoint blocks (they run concurrently without ordering), and no
    more steps than the circle method (block counts that do not halve evenly keep the circle method: npairs == 0).  Host function: no GPU."""
    import ctypes as C
    import numpy as np
    import tnb200
    lib = tnb200.load()
    used = 0
    for groups in (2, 4):
        for nb in list(range(2, 71, 2)) + [96, 128, 192, 256]:
            cap = nb * (nb - 1) // 2
            buf = (C.c_int32 * (5 * cap))()
            n = C.c_int64()
            assert lib.tn_svd_split_schedule(nb, groups, buf, cap, C.byref(n)) == 0
            if n.value == 0:
                continue
            used += 1
            assert n.value == cap, (nb, groups, n.value)
            a = np.frombuffer(buf, dtype=np.int32).reshape(cap, 5)
            pairs = {(int(p), int(q)) for p, q in a[:, 3:]}
            assert len(pairs) == cap and all(0 <= p < q < nb for p, q in pairs)
            depth = 0
            for ph in np.unique(a[:, 0]):
                rows = a[a[:, 0] == ph]
                blocks_of_task = []
                for t in np.unique(rows[:, 1]):
                    rt = rows[rows[:, 1] == t]
                    for st in np.unique(rt[:, 2]):
                        blk = rt[rt[:, 2] == st][:, 3:].reshape(-1)
                        assert len(set(blk.tolist())) == len(blk), "a step must rotate disjoint pairs"
                    blocks_of_task.append(set(rt[:, 3:].reshape(-1).tolist()))
                for i in range(len(blocks_of_task)):
                    for j in range(i):
                        assert not (blocks_of_task[i] & blocks_of_task[j]), "concurrent tasks must own disjoint blocks"
                assert 2 <= len(blocks_of_task) <= groups
                depth += max(int(rows[rows[:, 1] == t][:, 2].max()) + 1 for t in np.unique(rows[:, 1]))
            assert depth == nb - 1
    assert used >= 20
