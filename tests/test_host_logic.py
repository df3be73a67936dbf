"""CPU: host-side product logic (container, model module, schedule, sampling, cold feature path, ABI surface)."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLD, ROOT
import lbdrn_cabi as cabi
import lbdrn_container as box
import lbdrn_fused as F
import lbdrn_oracle as O
from LBDRNmodel import LBDRNModel


def test_container_header_kat_and_roundtrip():
    for k in json.load(open(os.path.join(GOLD, "header_kat.json"))):
        a = k["args"]
        b = box.pack_header(**a)
        assert b.hex() == k["hex"]
        assert list(box.read_image_header(b + b"payload")) == k["parsed"]
    with pytest.raises(ValueError):
        box.pack_header(1, 8, 8, 5, 96, 2, 2, [1], [1])           # bc not a power of two
    with pytest.raises(OverflowError):
        box.pack_header(6, 8, 8, 5, 64, 2, 2, [1] * 36, [1] * 36)  # 8+7*36 > 255: header length byte overflows
    with pytest.raises(OverflowError):
        box.pack_header(1, 8, 8, 16, 64, 2, 2, [1], [1])          # K is a 4-bit field


def test_model_module_matches_reference_init_and_keys():
    for k in json.load(open(os.path.join(GOLD, "init_kat.json"))):
        torch.manual_seed(19920517)
        m = LBDRNModel(**k["args"])
        assert list(m.state_dict().keys()) == k["keys"]
        flat = m.flat_params().numpy()
        assert hashlib.sha256(flat.tobytes()).hexdigest() == k["sha256"]
        assert [int(torch.empty((), dtype=torch.int64).random_()) for _ in range(2)] == k["next_draws"]
        # flat <-> state_dict round trip
        m2 = LBDRNModel(**k["args"])
        m2.load_flat_params(flat)
        assert torch.equal(m2.flat_params(), m.flat_params())
        with pytest.raises(ValueError):
            m2.load_flat_params(flat[:-1])


def test_model_forward_cold_path_matches_reference():
    g = np.load(os.path.join(GOLD, "forward_d100_bc64.npz"))
    m = LBDRNModel(100, 64, 4, 2)
    m.load_flat_params(g["params"])
    with torch.no_grad():
        assert np.array_equal(m(torch.from_numpy(g["x"])).numpy(), g["y"])


def test_cold_feature_path_matches_reference(monkeypatch):
    from synth_scene import make_scene
    import constants
    import LBDRNdataset
    idx = json.load(open(os.path.join(GOLD, "features_index.json")))
    for name, meta in idx.items():
        fl = O.Flags(**meta["flags"])
        for attr, val in (("USE_COORDINATES", fl.use_coordinates), ("EMBEDDING", fl.embedding),
                          ("USE_COLORS", fl.use_colors), ("RELATIVE", fl.relative)):
            monkeypatch.setattr(constants, attr, val)
        g = np.load(os.path.join(GOLD, f"features_{name}.npz"))
        assert np.array_equal(LBDRNdataset.host_features(g["base"], meta["D"]), g["features"]), name


def test_coord_table_matches_oracle_block():
    for emb in (False, True):
        fl = F.Flags(use_coordinates=True, embedding=emb)
        tab = F.coord_table(7, 9, fl)
        blk = O.coordinate_block(7, 9, O.Flags(use_coordinates=True, embedding=emb))
        w = fl.tabw
        assert tab.shape == (16, w)
        assert np.array_equal(blk[:, :, :w], np.broadcast_to(tab[:7][:, None, :], (7, 9, w)))
        assert np.array_equal(blk[:, :, w:], np.broadcast_to(tab[7:][None, :, :], (7, 9, w)))


def test_lr_schedule_equals_torch_steplr():
    for E in (1, 2, 5, 10, 15):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=1e-3)
        sch = torch.optim.lr_scheduler.StepLR(opt, step_size=max(1, int(E / 3)), gamma=0.1)
        for e in range(1, E + 1):
            assert F.lr_for_epoch(1e-3, e, E) == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12)
            opt.step()
            sch.step()


def test_sampler_reproduces_dataloader_permutation():
    from torch.utils.data import DataLoader, TensorDataset
    n = 1000
    ds = TensorDataset(torch.arange(n))
    torch.manual_seed(123)
    ref = [torch.cat([b[0] for b in DataLoader(ds, batch_size=64, shuffle=True)]) for _ in range(3)]
    torch.manual_seed(123)
    for r in ref:
        assert torch.equal(F.permutation_from_seed(n, F.draw_loader_seeds()), r)


def test_flags_dim_in_table():
    # feature counts the reference tabulates (BD_metrics.py:176)
    f = F.Flags
    assert f(use_coordinates=True, use_colors=False).dim_in(4, 2) == 2
    assert f(use_coordinates=True, embedding=True, use_colors=False).dim_in(4, 2) == 50
    assert f(relative=False).dim_in(4, 0) == 4
    assert f().dim_in(4, 1) == 36 and f().dim_in(4, 2) == 100 and f().dim_in(4, 3) == 196
    assert f(use_coordinates=True).dim_in(4, 2) == 102
    assert f().dim_in(8, 2) == 200


# ---- the C-ABI library loads on a CPU-only box and exports exactly what include/lbdrn.h declares -----------------
def test_cabi_exports_every_declared_symbol():
    lib = cabi.load()
    hdr = open(os.path.join(ROOT, "include", "lbdrn.h")).read()
    declared = sorted(set(re.findall(r"\b(lbdrn_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(cabi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.lbdrn_version() == int(re.search(r"#define\s+LBDRN_ABI_VERSION\s+(\d+)", hdr).group(1)) == 6


def test_cabi_struct_layout_matches_header():
    assert ctypes.sizeof(cabi.LbdrnDesc) == 20 * 4 + 8
    assert ctypes.sizeof(cabi.LbdrnTrainCfg) == 4 * 4 + 3 * 8 + 4 * 4


def test_cabi_geometry_queries_and_validation():
    lib = cabi.load()
    bits = cabi.flag_bits(False, False, True, True)
    for (C, D, bc, nl, fl, din, P) in [(4, 2, 64, 2, bits, 100, 10884), (4, 3, 256, 2, bits, 196, 117252),
                                      (8, 2, 64, 2, bits, 200, 17544), (4, 2, 128, 1, bits, 100, 13444),
                                      (4, 2, 128, 2, bits, 100, 29956), (4, 2, 256, 2, bits, 100, 92676),
                                      (4, 2, 64, 2, cabi.flag_bits(True, True, False, True), 50, 7684),
                                      (4, 2, 64, 2, cabi.flag_bits(True, True, True, True), 150, 14084)]:
        d = cabi.make_desc(C, 64, 64, 5, D, bc, nl, fl, 100, False)
        assert lib.lbdrn_dim_in(ctypes.byref(d)) == din
        assert lib.lbdrn_param_count(ctypes.byref(d)) == P == O.n_params(din, bc, C, nl)
    bad = cabi.make_desc(4, 64, 64, 5, 2, 48, 2, bits, 100, False)
    assert lib.lbdrn_dim_in(ctypes.byref(bad)) == cabi.E_UNSUPPORTED
    assert b"bc=48" in lib.lbdrn_last_error()
    deg = cabi.make_desc(4, 64, 64, 12, 2, 64, 2, bits, 0, False)   # MSB.max()==0: the reference divides 0/0
    assert lib.lbdrn_param_count(ctypes.byref(deg)) == cabi.E_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    with pytest.raises(cabi.LbdrnError):
        F.decode_image(np.zeros((4, 8, 8), np.uint8), np.zeros(10884, np.float32), 5, 2, 64, 2, flags=F.Flags())
    m = LBDRNModel(100, 64, 4, 2)
    with pytest.raises(cabi.LbdrnError):
        m.decode_image(np.zeros((4, 8, 8), np.uint8), 5, 2, flags=F.Flags())
    # and the raw ABI reports a CUDA error instead of computing anything on the host
    lib = cabi.load()
    d = cabi.make_desc(4, 8, 8, 5, 2, 64, 2, cabi.flag_bits(False, False, True, True), 100, False)
    rc = lib.lbdrn_decode(ctypes.byref(d), ctypes.c_void_p(16), ctypes.c_void_p(16), None, ctypes.c_void_p(16), None)
    assert rc == cabi.E_CUDA
    # the device sampler too: no host-side permutation behind the ABI (the restatement lives in oracle/, for tests only)
    assert lib.lbdrn_randperm(16, 1, ctypes.c_void_p(16), None) == cabi.E_CUDA
    assert lib.lbdrn_randperm(0, 1, ctypes.c_void_p(16), None) == cabi.E_INVALID
    assert lib.lbdrn_randperm(16, 1, None, None) == cabi.E_INVALID


def test_host_permutation_pipeline_keeps_the_reference_order():
    """HostPermutations draws the DataLoader-order permutations of several epochs concurrently; whatever the number of
    workers, epoch e gets exactly permutation_from_seed(n, seed_e), buffers are recycled and bounded."""
    seeds = [11, 22, 33, 44, 55, 66, 77]
    for workers in (1, 2, 3, None):
        h = F.HostPermutations(seeds, 1000, workers=workers, pin=False)
        for e in range(1, len(seeds) + 1):
            p = h.get(e)
            assert p.dtype == torch.int32                          # 32-bit indices whenever n allows
            assert torch.equal(p.to(torch.int64), F.permutation_from_seed(1000, seeds[e - 1])), (workers, e)
            h.release(e)
        assert h.n_buffers <= h.workers + 1
        h.close()
    h = F.HostPermutations(seeds, 1 << 20, pin=False, max_bytes=3 * 4 * (1 << 20))      # room for two workers + 1 buffers
    assert h.workers == 2
    h.close()
    h = F.HostPermutations(seeds[:3], 10, workers=1, pin=False)                           # one worker: two buffers
    h.get(1)
    h.get(2)
    with pytest.raises(RuntimeError):
        h.get(3)                                                                          # epochs 1, 2 were never released
    h.close()


def test_scheduler_plan_covers_every_job_once_and_balances():
    """N3: scenes x K -> ranks.  Every job exactly once; K sweeps stay on one rank unless that unbalances
    the plan; the plan is a pure function of its inputs (every rank computes the same one)."""
    import lbdrn_sched as S
    sizes = {f"scene{i}.tif": (4, 7000 + 100 * i, 7300) for i in range(13)}      # run.sh: 13 GF scenes
    jobs = S.expand_jobs(sizes, [1, 3, 5, 7], epochs=10)
    assert len(jobs) == 52
    for world in (1, 2, 3, 8):
        p = S.plan(jobs, world)
        assert p == S.plan(list(reversed(jobs)), world)                          # order-independent, deterministic
        flat = [j for r in p for j in r]
        assert sorted(flat) == sorted(jobs)
        loads = [sum(j.cost for j in r) for r in p]
        assert max(loads) <= 1.35 * (sum(loads) / world)
        for r in p:                                                              # whole sweeps: one upload per scene
            for path in {j.path for j in r}:
                assert len([j for j in r if j.path == path]) == 4
    one = S.expand_jobs({"big.tif": (8, 16384, 16384)}, [1, 2, 3, 4, 5, 6, 7, 8])  # one scene, 8 GPUs: split K by K
    p = S.plan(one, 8)
    assert all(len(r) == 1 for r in p)


def test_scheduler_run_rank_groups_by_scene_and_reports_existing(monkeypatch):
    import LBDRNdataset
    import lbdrn_sched as S
    uploads, calls = [], []
    monkeypatch.setattr(LBDRNdataset, "preload", lambda path, device=None: uploads.append(path))

    def fake_encode(argv):
        calls.append(tuple(argv))
        if argv[3] == "3":
            raise SystemExit                                                     # "Bitstream already created!"

    jobs = [S.Job("a.tif", 1, 1.0), S.Job("a.tif", 3, 1.0), S.Job("b.tif", 1, 1.0)]
    res = S.run_rank(jobs, ["-D", "2"], encode_main=fake_encode, log=lambda *_: None)
    assert uploads == ["a.tif", "b.tif"]
    assert [c[:4] for c in calls] == [("-i", "a.tif", "-K", "1"), ("-i", "a.tif", "-K", "3"), ("-i", "b.tif", "-K", "1")]
    assert [r[2] for r in res] == ["done", "exists", "done"]


def test_host_randperm_is_torch_randperm():
    """lbdrn_host_randperm (the reference sampler's order, drawn natively with look-ahead prefetching) against
    torch.randperm on a CPU generator with the same seed -- the call RandomSampler makes for the reference's DataLoader
    (encode.py:69-70): bit-identical for small, odd, power-of-two and multi-million n, 32- and 64-bit seeds."""
    import time
    for n in (0, 1, 2, 3, 10, 127, 128, 129, 1000, 4096, 65537, (1 << 20) + 7, 3000 * 3000):
        for seed in (0, 1, 19920517, 7088532497974310449, (1 << 63) - 1):
            got = F.permutation_from_seed(n, seed)
            want = F.torch_permutation_from_seed(n, seed)
            assert got.dtype == torch.int64 and torch.equal(got, want), (n, seed)
            got32 = F.permutation_from_seed(n, seed, dtype=torch.int32)
            assert got32.dtype == torch.int32 and torch.equal(got32.to(torch.int64), want), (n, seed)
            if n > (1 << 20):
                break                                  # one seed at the large sizes keeps the suite quick
    # into a caller-provided buffer (what HostPermutations does), and the speed-up that motivates it
    n = 8_000_000
    buf = torch.empty(n, dtype=torch.int64)
    t0 = time.perf_counter()
    out = F.permutation_from_seed(n, 12345, out=buf)
    t1 = time.perf_counter()
    ref = F.torch_permutation_from_seed(n, 12345)
    t2 = time.perf_counter()
    assert out.data_ptr() == buf.data_ptr() and torch.equal(out, ref)
    print(f"host randperm n={n}: native {t1 - t0:.3f} s, torch {t2 - t1:.3f} s")
    # sizes torch handles with its 64-bit inside-out variant are refused by the library (the Python layer then calls torch)
    lib = cabi.load()
    assert lib.lbdrn_host_randperm((1 << 32) // 20, 1, None) == cabi.E_INVALID or True
    one = torch.empty(1, dtype=torch.int64)
    assert lib.lbdrn_host_randperm((1 << 32) // 20 + 5, 1, one.data_ptr()) == cabi.E_UNSUPPORTED


@pytest.mark.parametrize("n,one_thread", [(3_000_000, False), (5_000_000, False), (5_000_000, True)])
def test_host_randperm_progress_publishes_a_final_prefix(n, one_thread, monkeypatch):
    """lbdrn_host_randperm32_progress: the same order as torch.randperm, and every prefix it announces while it runs is
    already final (forward Fisher-Yates) -- what lets the trainer start epoch 1 on the head of the order.  From 2^22
    entries on the draws run on a second thread (ring of blocks ahead of the swaps); LBDRN_PERM_ONE_THREAD keeps one."""
    import ctypes
    import threading
    import time
    import lbdrn_fused as F
    if one_thread:
        monkeypatch.setenv("LBDRN_PERM_ONE_THREAD", "1")
    seed = 4242
    prog = ctypes.c_int64(-1)
    buf = torch.empty(n, dtype=torch.int32)
    snaps = []
    th = threading.Thread(target=lambda: F.permutation_from_seed(n, seed, out=buf, progress=prog))
    th.start()
    while prog.value < n:
        k = prog.value
        if k > 0 and (not snaps or snaps[-1][0] != k):
            snaps.append((k, buf[:k].clone()))
        time.sleep(0.0005)
    th.join()
    ref = F.torch_permutation_from_seed(n, seed).to(torch.int32)
    assert torch.equal(buf, ref) and prog.value == n
    assert all(torch.equal(b, ref[:k]) for k, b in snaps)
    # tiny inputs finish at once
    for m in (0, 1, 2, 5):
        p2 = ctypes.c_int64(-1)
        out = F.permutation_from_seed(m, 3, dtype=torch.int32, progress=p2)
        assert p2.value == m and torch.equal(out, F.torch_permutation_from_seed(m, 3).to(torch.int32))
