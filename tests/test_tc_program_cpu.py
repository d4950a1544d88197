"""Host-built tile program of the tcgen05 chain kernel (`csrc/ren_tc.cu`: `build_forward` / `build_dgrad` /
`build_ops_impl`), checked without a GPU: the stage ops the MMA issuer walks, the weight-pack jobs the ops' bytes come
from, and the mbarrier hand-offs between the MMA issuer and the epilogue warps must stay consistent -- one unpaired
wait deadlocks the persistent kernel.  The program is dumped by the internal export `ren_tc_debug_program`.

What is emulated per tile, for the forward program with and without the bias K step and for the dgrad program:
  * ops and pack jobs both tile the packed-weight buffer contiguously, in the same order;
  * every K step of every op multiplies A features by the weight columns that hold exactly those features
    (TMEM column arithmetic of the epilogue's in-place operand write-back, shared-memory encoding segments);
  * every (output row, input column) of every layer's weight matrix is covered exactly once;
  * an accumulator range is overwritten before it is accumulated into;
  * waits on `bar_a[h]` equal the epilogue's arrivals, commits on `bar_acc[h]` equal the epilogue's waits;
  * a bias K step reads the layer's bias vector (state-dict offset) with the layer's log2-domain scale.
"""
import ctypes as C
import struct

import pytest

import __graft_entry__ as G

OP_ONES = 0x400
LOG2E_100 = 144.26950408889634
EPI_F_HIDDEN, EPI_F_LINEAR, EPI_F_LAST, EPI_D_ACT, EPI_D_LINEAR = range(5)


@pytest.fixture(scope="module")
def built_lib():
    return G.build()


def dump(lib, fdesc, rdesc, mode, bias_mma):
    steps = (C.c_int32 * (12 * 32))()
    ops = (C.c_uint32 * (8 * 128))()
    jobs = (C.c_int64 * (12 * 64))()
    n_steps, n_ops, n_jobs, pack_bytes = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    fn = lib.ren_tc_debug_program
    fn.restype = C.c_int
    rc = fn(C.byref(fdesc), C.byref(rdesc), C.c_int32(mode), C.c_int32(bias_mma), steps, C.byref(n_steps), ops,
            C.byref(n_ops), jobs, C.byref(n_jobs), C.byref(pack_bytes))
    assert rc == 0, rc
    keys_s = ["N", "acc_col", "a_col", "a_k", "seg", "epi", "out_mat", "act_mat", "write_a", "bias_idx", "r1sh",
              "bias_mma"]
    S = [dict(zip(keys_s, steps[12 * i:12 * i + 12])) for i in range(n_steps.value)]
    keys_o = ["pack_off", "bytes", "d_col", "a_addr", "a_lo_delta", "flags", "pad0", "pad1"]
    O = [dict(zip(keys_o, ops[8 * i:8 * i + 8])) for i in range(n_ops.value)]
    keys_j = ["dst_off", "mode", "n0", "Ntrue", "n_ksteps", "k0", "Ktrue", "rows", "w_off", "ldw", "col0", "scale"]
    J = [dict(zip(keys_j, jobs[12 * i:12 * i + 12])) for i in range(n_jobs.value)]
    for j in J:
        j["scale"] = struct.unpack("f", struct.pack("I", j["scale"]))[0]
    return S, O, J, pack_bytes.value


def split_ops_by_step(S, O, mode):
    """Ops belong to steps in order; a step ends with the op that commits its last accumulator half."""
    out, i = [], 0
    for st in S:
        last = st["N"] >> 7
        cur = []
        while True:
            o = O[i]
            i += 1
            cur.append(o)
            if (o["flags"] >> 8) & 3 == last:
                break
        out.append(cur)
    assert i == len(O)
    return out


def check_program(lib, fdesc, rdesc, mode, bias_mma, layout, C_dim):
    S, O, J, pack_bytes = dump(lib, fdesc, rdesc, mode, bias_mma)
    assert len(O) <= 128 and len(J) <= 64
    wjobs = [j for j in J if j["mode"] in (0, 1)]
    # ---- both tilings of the pack buffer are contiguous and identical in extent
    off = 0
    for o in O:
        assert o["pack_off"] == off and o["bytes"] > 0 and o["bytes"] <= 32768 and o["bytes"] % 16 == 0
        off += o["bytes"]
    assert off == pack_bytes
    off = 0
    for j in wjobs:
        assert j["dst_off"] == off and j["rows"] in (128, 256)
        off += j["n_ksteps"] * j["rows"] * 64
    assert off == pack_bytes

    def job_of(byte_off):
        for j in wjobs:
            size = j["n_ksteps"] * j["rows"] * 64
            if j["dst_off"] <= byte_off < j["dst_off"] + size:
                return j
        raise AssertionError(byte_off)

    # ---- per step: operand pairing, coverage, accumulator initialisation
    by_step = split_ops_by_step(S, O, mode)
    waits = {1: 0, 2: 0, 3: 0}
    commits = {1: 0, 2: 0}
    n_bias_ops = 0
    for g, (st, ops) in enumerate(zip(S, by_step)):
        nh = st["N"] >> 7
        initialised = set()           # accumulator halves (0 / 1) already overwritten in this step
        cover = {}                    # (weight w_off, n, k) -> count
        bias_seen = False
        for oi, o in enumerate(ops):
            fl = o["flags"]
            cnt, wide, a_smem, fresh = fl & 7, bool(fl & 8), bool(fl & 16), bool(fl & 32)
            wt, ca = (fl >> 6) & 3, (fl >> 8) & 3
            if wt:
                waits[wt] += 1
            if ca:
                commits[ca] += 1
            N_op = 256 if wide else 128
            assert o["d_col"] >= st["acc_col"] and o["d_col"] + N_op <= st["acc_col"] + st["N"]
            halves = {(o["d_col"] - st["acc_col"]) // 128 + i for i in range(N_op // 128)}
            if fresh:
                assert not (halves & initialised), "accumulator overwritten twice"
                initialised |= halves
            else:
                assert halves <= initialised, "accumulate into an uninitialised accumulator"
            j = job_of(o["pack_off"])
            assert j["rows"] == N_op
            chunk = N_op * 64
            assert (o["pack_off"] - j["dst_off"]) % chunk == 0 and o["bytes"] == cnt * chunk
            c0 = (o["pack_off"] - j["dst_off"]) // chunk
            assert c0 + cnt <= j["n_ksteps"]
            assert j["n0"] == o["d_col"] - st["acc_col"]
            if fl & OP_ONES:
                # bias K step: first op of the step, overwrites the whole accumulator, B = the bias vector
                n_bias_ops += 1
                assert mode == 0 and st["bias_mma"] == 1 and oi == 0 and fresh and a_smem and cnt == 1
                assert N_op == st["N"] and o["d_col"] == st["acc_col"]
                w_off, b_off, out_f, in_f = layout[st["bias_idx"]]
                assert (j["w_off"], j["ldw"], j["Ktrue"], j["k0"], j["col0"], j["n_ksteps"]) == (b_off, 1, 1, 0, 0, 1)
                assert j["Ntrue"] == out_f and abs(j["scale"] - LOG2E_100) < 1e-3
                bias_seen = True
                continue
            for kl in range(cnt):
                # A features of this K step (index into the layer's concatenated input [main | segment])
                if a_smem:
                    base = 0 if st["seg"] == 1 else 32768            # SM_P_HI / SM_D_HI
                    ks = ((o["a_addr"] << 4) - base) // 4096 + kl
                    assert ((o["a_addr"] << 4) - base) % 4096 == 0
                    feat = ("seg", 16 * ks)
                    assert o["a_lo_delta"] << 4 == (16384 if st["seg"] == 1 else 8192)
                else:
                    # the epilogue wrote features n0 + 8 j8 .. + 7 of the previous step to columns
                    # a_col + n0 + 4 j8 (hi) and + 16 (lo): K step ks of half hh starts at 32 (ks >> 1) + 8 (ks & 1)
                    col = o["a_addr"] + 32 * (kl >> 1) + 8 * (kl & 1) - st["a_col"]
                    hh, rem = divmod(col, 128)
                    ks2, r8 = divmod(rem, 32)
                    assert r8 in (0, 8) and o["a_lo_delta"] == 16
                    feat = ("main", 128 * hh + 16 * (2 * ks2 + (1 if r8 else 0)))
                    assert wt in (0, hh + 1) if kl == 0 else True
                # weight columns the B chunk holds
                kcol = j["k0"] + 16 * (c0 + kl)
                if feat[0] == "main":
                    assert j["col0"] == 0 and kcol == feat[1], (g, o, j, feat)
                else:
                    assert j["k0"] == 0 and kcol == feat[1], (g, o, j, feat)
                for n in range(j["n0"], j["n0"] + N_op):
                    key = (j["w_off"], j["mode"], j["col0"], n, kcol)
                    cover[key] = cover.get(key, 0) + 1
        assert initialised == set(range(nh))
        assert all(v == 1 for v in cover.values())
        if mode == 0:
            assert bias_seen == bool(st["bias_mma"])
        # K coverage: a_k main features (+ the segment's K steps), every output half
        n_ksteps = st["a_k"] // 16 + (4 if st["seg"] == 1 else (2 if st["seg"] == 2 else 0))
        assert len(cover) == n_ksteps * st["N"]
    # ---- hand-offs: the MMA issuer waits on bar_a[h] once per epilogue arrival, commits bar_acc[h] once per wait
    arr0 = sum(1 for st in S if st["write_a"]) + (1 if mode == 1 else 0)      # dgrad: + the tile prologue
    arr1 = sum(1 for st in S if st["write_a"] and st["N"] == 256)
    assert waits[1] == arr0 and waits[2] == arr1
    assert waits[3] == (1 if mode == 0 else 0)
    assert commits[1] == len(S) and commits[2] == sum(1 for st in S if st["N"] == 256)
    return S, O, J, n_bias_ops


@pytest.mark.parametrize("radiance_dim", [1, 3])
def test_tile_programs(built_lib, radiance_dim):
    from ren_b200 import _lib
    lib = C.CDLL(built_lib)
    f = _lib.make_field_desc(radiance_dim=radiance_dim)
    r = _lib.make_render_desc(near_plane=0.5, far_plane=4.5, render_step_size=4 / 128)
    layout = _lib.param_layout(f)
    S0, O0, J0, nb0 = check_program(lib, f, r, 0, 0, layout, radiance_dim)
    S1, O1, J1, nb1 = check_program(lib, f, r, 0, 1, layout, radiance_dim)
    check_program(lib, f, r, 1, 0, layout, radiance_dim)
    assert nb0 == 0 and nb1 == 8                       # the eight 256-wide softplus layers of the trunk
    assert len(O1) == len(O0) + 8 and len(J1) == len(J0) + 8
    # the bias K step changes nothing else: dropping the bias ops gives the plain program shifted by their bytes
    plain = [o for o in O1 if not o["flags"] & OP_ONES]
    shift = 0
    it = iter(O1)
    for o0, o1 in zip(O0, plain):
        while True:
            o = next(it)
            if o["flags"] & OP_ONES:
                shift += o["bytes"]
                moved = (o["flags"] >> 6) & 3
            else:
                break
        assert o1["pack_off"] == o0["pack_off"] + shift
        for k in ("bytes", "d_col", "a_addr", "a_lo_delta"):
            assert o1[k] == o0[k]
        # flags differ only where the bias op took over the step's first wait and the overwrite
        diff = o0["flags"] ^ o1["flags"]
        assert diff & ~((3 << 6) | 32) == 0
    # the layers that take the bias through the GEMM are the hidden trunk layers, and only in the forward program
    assert [s["bias_mma"] for s in S1] == [1] * 8 + [0, 0]
    assert [s["epi"] for s in S1] == [EPI_F_HIDDEN] * 8 + [EPI_F_LINEAR, EPI_F_LAST]


def test_no_per_lane_serialisation_loops_in_the_tcgen05_kernels(built_lib):
    """The single-thread roles (TMA producers, MMA issuers) are guarded by `elect.sync`: under `lane == 0` ptxas
    wraps every instruction with uniform-register operands (tcgen05.mma, tcgen05.commit, bulk copies) in an
    `ELECT ... BRA.U.ANY` loop over the active lanes -- 213 of them in ren_tc.o -- and the issuing thread becomes
    the bottleneck of the tensor pipe (forward -9 %, dgrad -6 %).  Guard against a
    regression by looking at the SASS of the objects the library was linked from."""
    import os
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    obj_dir = os.path.join(G.PKG_DIR, "build")
    for name in ("ren_tc.o", "ren_tc_wgrad.o", "ren_tc_raygrad.o"):
        path = os.path.join(obj_dir, name)
        if not os.path.exists(path):
            pytest.skip("objects of the in-tree build not present")
        sass = subprocess.run([cuobjdump, "-sass", path], capture_output=True, text=True, check=True).stdout
        assert "UTCHMMA" in sass, name                      # tcgen05.mma is what these kernels issue
        assert sass.count("BRA.U.ANY") == 0, f"{name}: per-lane serialisation loops around uniform-datapath instructions"


def test_schedule_model_reproduces_the_measured_operating_points():
    """`tools/chain_model.py`: with the measured per-layer epilogue / tensor / hand-off cycles
    the discrete-event model of the two-halves schedule lands on both measured tensor-pipe figures -- 52 % with the
    lane-0 issuer's MMA time (8.4 k cycles per layer), 60 % with the elect.sync issuer -- and ranks the re-tilings."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "chain_model", os.path.join(G.ROOT, "tools", "chain_model.py"))
    cm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cm)
    active = 5650.0                                    # measured tensor-pipe active cycles per 256-wide layer
    today = cm.period([0.5, 0.5], 6200.0, 6500.0, 700.0, 700.0)
    slow = cm.period([0.5, 0.5], 6200.0, 8400.0, 700.0, 700.0)
    assert abs(active / today - 0.600) < 0.01          # profiles/r02_tc_step_ncu_summary_elect.txt: 60.0 %
    assert abs(active / slow - 0.523) < 0.01           # profiles/r02_tc_step_ncu_summary.txt: 52.3 %
    quarters = cm.period([0.25] * 4, 6200.0, 6500.0, 700.0, 700.0)
    lopsided = cm.period([0.25, 0.75], 6200.0, 6500.0, 700.0, 700.0)
    assert quarters < 0.9 * today < lopsided           # four quarters help, a 64 | 192 split does not
