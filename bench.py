#!/usr/bin/env python
"""bench.py -- headline benchmark of the LBDRN hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-encode]
    torchrun ... bench.py --gpus N ...         (one rank per GPU; rank 0 prints ONE JSON line)

Metric (BASELINE.json): decode throughput in Mpix/s at K=5 D=2 bc64 nl2; encode s/scene reported beside it.
A "step" is one fused decode pass over one synthetic scene resident in HBM.  Workload at N=1: configs[1] of
BASELINE.json, the 4-band 12-bit 8192x8192 scene.  At N>1 the scene grows to N*8192 rows and is row-stripe
sharded (weak scaling: 8192 rows per GPU).  "Resident" includes what makes a stripe decodable: its rows, the scene's
global MSB maximum and the D halo rows of the neighbouring stripes are exchanged ONCE when the scene is made resident
(`StreamedStripeDecoder.preload`: one all-reduce + one halo swap, SURVEY 8e "one-off"), not per timed step.
`e2e` is the same metric through the public streaming API (`lbdrn_fused.StreamedDecoder` /
`lbdrn_dist.StreamedStripeDecoder`) with pinned HOST buffers: H2D of the base layer, the per-scene max reduction
[+ all-reduce + halo swap at N>1] and D2H of the reconstruction inside the timed region.
`--impl reference` times the reference's CPU implementation of the path: the oracle port (oracle/lbdrn_oracle.py,
a restatement pinned bit-exactly against the unmodified reference) on the box's host cores, on a bounded crop.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

C_, SIDE, BITS, K_, D_, BC, NL = 4, 8192, 12, 5, 2, 64, 2
DIM_IN = C_ * (2 * D_ + 1) ** 2
FLOP_PER_PX = 2 * (DIM_IN * BC + (NL - 1) * BC * BC + BC * C_)          # 21 504
TRAIN_FLOP_PER_PX = 3 * FLOP_PER_PX - 2 * DIM_IN * BC                    # 51 712
HBM_BYTES_PER_PX = C_ * 1 + C_ * 2                                       # u8 MSB in, u16 out
METRIC, UNIT = "decode_throughput_K5_D2_bc64_nl2", "Mpix/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(bf16_burst=j["bf16_tflops"], bf16_sustained=j["bf16_tflops_sustained"], hbm=j["hbm_gbs"],
                    source="MEASURED_PEAKS.json")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every ~2 ms on a thread (the timed
    region of a decode benchmark is only tens of ms, too short for `nvidia-smi -lms`)."""

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th, self.h = index, [], False, None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.h = None
        return self

    def _poll(self):
        nv = self.nv
        while not self.stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(mhz), int(reasons)))
            except Exception:
                pass
            time.sleep(0.002)

    def __exit__(self, *a):
        self.stop = True
        if self.th:
            self.th.join(timeout=2)

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        seen = 0
        for _, r in self.rows:
            seen |= r
        return dict(sm_mhz=statistics.median(m for m, _ in self.rows), sm_max_mhz=self.max_mhz,
                    reasons=sorted(k for k, bit in names.items() if seen & bit), samples=len(self.rows))


def bench_params(device):
    """Decode-only benchmark weights: reference init from the reference's seed, then the fpzip(prec=16) value map
    (low 16 bits cleared) -- what a decoder would load."""
    from LBDRNmodel import LBDRNModel
    torch.manual_seed(19920517)
    flat = LBDRNModel(DIM_IN, BC, C_, NL).flat_params()
    flat = (flat.view(torch.int32) & -65536).view(torch.float32)
    return flat.to(device)


KERNEL_PROFILE = os.path.join(ROOT, "profiles", "r2_decode_kernel_ncu.json")


def kernel_profile(has_tc):
    """What ncu measured for ONE launch of the dominant kernel at the bench workload: DRAM bytes, warp instructions, issue
    utilisation -- read from the committed capture of the CURRENT kernel (written by tools/ncu_summary.py --traffic-json
    from `ncu --set full`; no literals here).  None if absent or if this GPU runs the fp32 kernel."""
    if not has_tc or not os.path.exists(KERNEL_PROFILE):
        return None
    j = json.load(open(KERNEL_PROFILE))
    j["bytes_per_launch"] = j["dram_bytes_read"] + j["dram_bytes_write"]
    j["algorithmic_bytes"] = SIDE * SIDE * HBM_BYTES_PER_PX
    return j


def cpu_reference_note():
    """Verbatim-reference vs port on the build container's CPU (oracle/time_reference_cpu.py; the reference is Python and
    cannot travel to the GPU box, so the box times the port and this file relates the two)."""
    p = os.path.join(ROOT, "profiles", "r2_cpu_reference_verbatim.json")
    return json.load(open(p)) if os.path.exists(p) else None


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_decode_sample(side, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lbdrn_oracle as O
    from synth_scene import make_scene
    torch.set_num_threads(threads)
    img = make_scene(C_, side, side, BITS, seed=19920517)
    msb, _ = O.split_msb_lsb(img, K_)
    torch.manual_seed(19920517)
    params = O.unflatten_params(O.fpzip_value_map(O.flatten_params(O.init_params(DIM_IN, BC, C_, NL)), 16),
                                DIM_IN, BC, C_, NL)
    t0 = time.perf_counter()
    O.decode_image(msb, params, K_, D_)
    return time.perf_counter() - t0


REF_SIDE = 2048          # CPU sample: one 4 x 2048 x 2048 scene per step, whatever N is (per-pixel metric, linear in N)


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    probe = cpu_decode_sample(256, threads)
    side = REF_SIDE
    if probe / (256 * 256) * side * side * (args.steps + args.warmup) > 900.0:      # would not end within minutes
        side = 1024
    for _ in range(args.warmup):
        cpu_decode_sample(side, threads)
    ts = [cpu_decode_sample(side, threads) for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    v = side * side / t / 1e6
    sample = f"{C_}x{side}x{side} synthetic scene per step (same generator / weights / K, D, bc, nl as the GPU arm)"
    note = cpu_reference_note()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: synthetic 4-band 12-bit scene, K=5 D=2 bc64 nl2, decode; CPU sample {sample}"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "verbatim_reference": note},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def _timed(fn, reps, dev, world):
    """fn() queues one step on the current stream; returns ms per step (CUDA events, max over ranks)."""
    import torch.distributed as dist
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_ours(args):
    import ctypes
    import torch.distributed as dist
    import lbdrn_cabi as cabi
    import lbdrn_dist as LD
    import lbdrn_fused as F
    from LBDRNmodel import LBDRNModel
    from synth_scene import make_scene_torch

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = cabi.load()
    fl = F.Flags()
    pk = peaks()
    cur = torch.cuda.current_stream(dev)

    # ---- synthetic scene: this rank's 8192-row stripe of a (world*8192)-row scene -------------------------------
    H_total = SIDE * world
    own = make_scene_torch(C_, SIDE, SIDE, BITS, seed=19920517 + rank, device=dev)        # CHW uint16
    scene = F.DeviceScene.from_image(own, K_)
    del own
    params = bench_params(dev)

    # N > 1: the streamed stripe decoder in resident mode: preload() exchanges the global max and the halo rows once (they are
    # static input, like the planes), a step then queues this rank's kernels only.  The per-scene collectives are paid inside
    # the timed region of `e2e` below.
    sdec = None
    if world > 1:
        sdec = LD.StreamedStripeDecoder(H_total, SIDE, C_, D_, scene.msb.dtype, K_, BC, NL, params, fl, dev, sub_rows=SIDE)
        sdec.preload(scene.msb)
    out = torch.empty((C_, SIDE, SIDE), dtype=torch.uint16, device=dev)

    def decode_step():
        if world == 1:
            d = scene.desc(D_, BC, NL, fl, path=cabi.PATH_AUTO)
            cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(scene.msb), cabi.ptr(params), None, cabi.ptr(out),
                                        cabi.stream_ptr()))
        else:
            cur.wait_event(sdec.submit())

    for _ in range(max(3, args.warmup)):
        decode_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = lib.lbdrn_launch_count()
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(args.steps):
            decode_step()
            ev[i + 1].record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lib.lbdrn_launch_count() - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    npx_step = SIDE * SIDE * world
    value = npx_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- N > 1: the stripes the ranks just produced equal a ONE-GPU decode of the same rows -------------------------------
    # (the driver's GPU-test box has one GPU and skips tests/test_gpu_dist.py, so the scaling run carries its own check)
    stripe_check = None
    if world > 1:
        r0, r1 = LD.stripe_bounds(H_total, world, rank)
        mx = torch.tensor([scene.msb_max], dtype=torch.int64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        gmax = int(mx.item())
        parts, top = [], 0
        if rank > 0:                                   # the neighbour stripes are synthetic too: rebuild their edge rows here
            nb = F.DeviceScene.from_image(make_scene_torch(C_, SIDE, SIDE, BITS, seed=19920517 + rank - 1, device=dev), K_)
            parts.append(nb.msb[:, SIDE - D_:].clone())
            top = D_
            del nb
        parts.append(scene.msb)
        if rank + 1 < world:
            nb = F.DeviceScene.from_image(make_scene_torch(C_, SIDE, SIDE, BITS, seed=19920517 + rank + 1, device=dev), K_)
            parts.append(nb.msb[:, :D_].clone())
            del nb
        buf = torch.cat(parts, dim=1).contiguous()
        ref_out = torch.empty_like(buf, dtype=torch.uint16)
        d = cabi.make_desc(C_, H_total, SIDE, K_, D_, BC, NL, fl.bits(), gmax, scene.msb_u16, row0=r0, row1=r1,
                           buf_row0=r0 - top, buf_rows=buf.shape[1])
        cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(buf), cabi.ptr(params), None, cabi.ptr(ref_out), cabi.stream_ptr()))
        got = sdec.result(sdec.n - 1)
        same = bool(torch.equal(got.view(torch.int16), ref_out[:, top:top + SIDE].view(torch.int16)))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        chk = got.view(torch.int16).to(torch.int64).sum().reshape(1)
        dist.all_reduce(chk)
        stripe_check = {"bit_identical_to_single_gpu_decode_of_the_same_rows": bool(flag.item()), "ranks": world,
                        "rows_per_rank": SIDE, "global_msb_max": gmax, "checksum_all_ranks": int(chk.item())}
        del buf, ref_out, parts
        if not stripe_check["bit_identical_to_single_gpu_decode_of_the_same_rows"]:
            raise SystemExit(f"stripe-sharded decode differs from the single-GPU decode of the same rows: {stripe_check}")

    # ---- dominant kernel: measured alone with CUDA events on its stream (the decode kernel IS the step at N=1) ----
    d = scene.desc(D_, BC, NL, fl, path=cabi.PATH_AUTO)
    has_tc = bool(lib.lbdrn_has_tensor_path(ctypes.byref(d)))
    if world == 1:
        kern_avg_ms = sum(step_ms) / len(step_ms)
    else:                                              # local kernel on this rank's rows, no collectives
        def local_kernel():
            cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(scene.msb), cabi.ptr(params), None, cabi.ptr(out),
                                        cabi.stream_ptr()))
        kern_avg_ms = _timed(local_kernel, 5, dev, world)
    prof = kernel_profile(has_tc)
    tflops = SIDE * SIDE * FLOP_PER_PX / (kern_avg_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "achieved": tflops, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                "frac": tflops / pk["bf16_burst"],
                "traffic": prof["bytes_per_launch"] if prof else None, "traffic_detail": prof,
                "kernel": "tc_pipe_kernel<MUFU> (tcgen05 kind::f16 from TMA-staged bytes, fp16 hi+lo split activations, fp32 TMEM "
                          "accumulators x2 per warpgroup, software-pipelined)" if has_tc else "infer_fp32_kernel<64,8,4,true,DECODE> (fp32 FFMA)",
                "kernel_ms": kern_avg_ms,
                "peak_source": pk["source"] + " bf16 dense burst",
                "algorithmic_flop_per_pixel": FLOP_PER_PX,
                "hbm": {"achieved_gbs": SIDE * SIDE * HBM_BYTES_PER_PX / (kern_avg_ms * 1e-3) / 1e9, "peak_gbs": pk["hbm"],
                        "algorithmic_bytes_per_pixel": HBM_BYTES_PER_PX}}
    if prof:
        # the resource that actually binds this kernel: instruction issue on the CUDA cores -- the sines of the
        # reference's activation, their fp16 hi+lo split and the feature build.  thread-instructions per pixel =
        # smsp__inst_executed.sum * 32 / pixels of the committed capture of THIS kernel; ceiling = 148 SMs x 128 lanes x clock.
        instr_px = prof["thread_instr_per_pixel"]
        ceil_gpix = 148 * 128 * (clk.summary().get("sm_mhz") or 1965.0) * 1e6 / instr_px / 1e9
        roofline["issue"] = {"thread_instr_per_pixel": instr_px, "ceiling_gpix_s": ceil_gpix,
                             "achieved_gpix_s": SIDE * SIDE / (kern_avg_ms * 1e-3) / 1e9,
                             "frac": SIDE * SIDE / (kern_avg_ms * 1e-3) / 1e9 / ceil_gpix,
                             "source": os.path.relpath(KERNEL_PROFILE, ROOT)}

    # ---- e2e: public API, pinned host buffers, H2D + D2H inside the timed region ------------------------------------
    base_host = scene.msb.cpu().pin_memory()
    outs_host = [torch.empty((C_, SIDE, SIDE), dtype=torch.uint16).pin_memory() for _ in range(2)]
    params_host = params.cpu()
    e2e_steps = max(8, min(2 * args.steps, 32))     # a stream of scenes: the pipeline's fill / drain (about one scene) is amortised
    if world == 1:
        streamer = F.StreamedDecoder(C_, SIDE, SIDE, base_host.dtype, K_, D_, BC, NL, params_host, flags=fl)
        submit = lambda i: streamer.submit(base_host, outs_host[i % 2])
    else:
        streamer = LD.StreamedStripeDecoder(H_total, SIDE, C_, D_, scene.msb.dtype, K_, BC, NL, params_host, fl, dev)
        submit = lambda i: streamer.submit(base_host, outs_host[i % 2])
    submit(0).synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    # a stream of scenes through the public streaming API: every scene is uploaded from pinned host memory, decoded (device-side
    # max [+ all-reduce + halo swap], stripe kernels) and downloaded to pinned host memory; scene i+1's upload overlaps scene i
    tickets = []
    for i in range(e2e_steps):
        tickets.append(submit(i))
        if i >= 1:
            tickets[i - 1].synchronize()
    tickets[-1].synchronize()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d_b = int(base_host.numel() * base_host.element_size())
    d2h_b = int(outs_host[0].numel() * 2)
    # host-side ceiling: the same bytes copied H2D and D2H on two streams by every rank at once, no kernels
    s_a, s_b = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dev_in, dev_out = torch.empty_like(scene.msb), torch.empty((C_, SIDE, SIDE), dtype=torch.uint16, device=dev)

    def copy_pair():
        with torch.cuda.stream(s_a):
            dev_in.copy_(base_host, non_blocking=True)
        with torch.cuda.stream(s_b):
            outs_host[0].copy_(dev_out, non_blocking=True)
    copy_pair()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        copy_pair()
    torch.cuda.synchronize()
    cp_s = (time.perf_counter() - t0) / 4
    if world > 1:
        t = torch.tensor([cp_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cp_s = float(t.item())
    e2e_v = npx_step * e2e_steps / e2e_s / 1e6
    e2e = {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d_b * world + int(params_host.numel()) * 4 * world,
           "d2h_bytes_per_step": d2h_b * world, "steps": e2e_steps,
           "host_copy_ceiling": {"Mpix_s": npx_step / cp_s / 1e6, "h2d_GBs_per_gpu": h2d_b / cp_s / 1e9,
                                 "d2h_GBs_per_gpu": d2h_b / cp_s / 1e9,
                                 "what": "the step's H2D and D2H bytes copied concurrently on two streams by all ranks at once, no kernels"},
           "frac_of_host_copy_ceiling": e2e_v / (npx_step / cp_s / 1e6),
           "api": ("lbdrn_fused.StreamedDecoder.submit/wait" if world == 1 else "lbdrn_dist.StreamedStripeDecoder.submit per rank") +
                  " (pinned host in/out per scene; H2D | device-side max" + ("" if world == 1 else " + all-reduce + halo swap") +
                  " | stripe kernels | D2H on three streams, two buffer slots so consecutive scenes overlap)"}
    del base_host, outs_host, streamer, dev_in, dev_out

    # ---- encode s/scene (10 epochs, bs 8192, per-epoch eval + best-epoch select), scene-per-GPU replicas -----------
    encode = None
    if not args.no_encode:
        def run_encode(sampler):
            torch.manual_seed(19920517)
            model = LBDRNModel(DIM_IN, BC, C_, NL)
            tr = F.FusedTrainer(model, scene, D_, 1e-3, 8192, args.encode_epochs, flags=fl, sampler=sampler)
            torch.cuda.synchronize()
            with ClockSampler(local) as enc_clk:
                t0 = time.perf_counter()
                res = tr.run()
                torch.cuda.synchronize()
                enc_s = time.perf_counter() - t0
            tr.close()
            if world > 1:
                t = torch.tensor([enc_s], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                enc_s = float(t.item())
            return enc_s, res, enc_clk.summary()

        # the first encode of a process also pays the lazy load of the training / evaluation modules and the first device
        # allocations (0.05-0.2 s): timed and reported as `first_run_s_per_scene`; `s_per_scene` is the second, like the
        # decode leg's warm-up steps (a scheduler encodes many scenes per process)
        cold_s, _, _ = run_encode(args.sampler)
        enc_s, res, enc_clocks = run_encode(args.sampler)
        n_steps = len(res["losses"])
        flop = SIDE * SIDE * args.encode_epochs * (TRAIN_FLOP_PER_PX + FLOP_PER_PX)
        encode = {"s_per_scene": enc_s, "scenes_per_s_all_gpus": world / enc_s, "mode": "scene-per-GPU replicas",
                  "epochs": args.encode_epochs, "batch_size": 8192, "optimizer_steps": n_steps,
                  "us_per_step_incl_eval": enc_s / n_steps * 1e6, "sampler": args.sampler,
                  "final_val_mse": res["val_mse"][-1] if res["val_mse"] else None, "best_epoch": res["best_epoch"],
                  "tensor_roofline_frac": flop / enc_s / 1e12 / pk["bf16_sustained"], "clocks": enc_clocks,
                  "first_run_s_per_scene": cold_s,
                  "excludes": "GDAL read/write, JPEG-2000 base layer, fpzip (host, unchanged)"}
        if world == 1 and not args.no_extra:
            # the parity-pinned mode of encode.py's default (--sampler reference: the DataLoader's own permutations, drawn on
            # the host ahead of the device) next to the device sampler timed above
            other = "reference" if args.sampler == "device" else "device"
            o_s, o_res, _ = run_encode(other)
            encode[f"{other}_sampler_s_per_scene"] = o_s
            encode[f"{other}_sampler_final_val_mse"] = o_res["val_mse"][-1] if o_res["val_mse"] else None
            encode["sampler_note"] = ("device: lbdrn_randperm orders (quality pinned per seed against the oracle trained on the same "
                                      "orders, tests/test_gpu_train.py::test_device_sampler_encode_quality_is_pinned); reference: "
                                      "the reference DataLoader's exact batches (torch's CPU randperm restated natively, drawn by host threads "
                                      "ahead of the device; epoch 1 trained in slices as its order is finalised)")

    # ---- encode, data-parallel mode (N > 1): every rank holds the SAME scene and takes 1/N of each batch; one NCCL
    # all-reduce of the P+1 gradient floats per step (lbdrn_dist.DataParallelTrainer).  Reported as measured next to the
    # scene-per-GPU mode above; latency-bound at bs = 8192 (three launches + a collective per step), expected <= 1x.
    if world > 1 and not args.no_encode and encode is not None:
        shared = F.DeviceScene.from_image(make_scene_torch(C_, SIDE, SIDE, BITS, seed=19920517, device=dev), K_)
        torch.manual_seed(19920517)
        tr = F.FusedTrainer(LBDRNModel(DIM_IN, BC, C_, NL), shared, D_, 1e-3, 8192, 1, flags=fl, sampler="device")
        tr.begin()
        dp = LD.DataParallelTrainer(tr)
        dp_steps = 256
        perm = torch.empty(SIDE * SIDE, dtype=torch.int64, device=dev)
        cabi.check(lib.lbdrn_randperm(SIDE * SIDE, 19920517, cabi.ptr(perm), cabi.stream_ptr()))   # same order on every rank
        dp.train_epoch(perm[:16 * 8192], 1e-3)
        torch.cuda.synchronize()
        dist.barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        dp_losses = dp.train_epoch(perm[16 * 8192:(16 + dp_steps) * 8192], 1e-3)
        d1.record()
        torch.cuda.synchronize()
        t = torch.tensor([d0.elapsed_time(d1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dp_us = float(t.item()) * 1e3 / dp_steps
        fused_us = encode["us_per_step_incl_eval"]
        encode["data_parallel"] = {"us_per_step": dp_us, "steps_timed": dp_steps, "n_gpus": world,
                                   "allreduce_bytes_per_step": 4 * (int(params.numel()) + 1),
                                   "speedup_vs_one_gpu_fused_step": fused_us / dp_us,
                                   "last_loss": float(dp_losses[-1].item()),
                                   "note": "per step: gradient kernel on 1/N of the batch, NCCL all-reduce, Adam kernel"}
        tr.close()
        del shared, perm

    # ---- the other BASELINE.json configs (explanatory lines, not the headline) -------------------------------------------
    other = None
    if not args.no_extra:
        other = []

        def trunc16(flat):
            return (flat.view(torch.int32) & -65536).view(torch.float32)

        def timed_decode(D, bc, flags, label, reps=3, relu=False, gains=None):
            torch.manual_seed(19920517)
            dim_in = flags.dim_in(C_, D)
            m = LBDRNModel(dim_in, bc, C_, NL, activation=torch.nn.ReLU() if relu else None)
            if gains:                                  # spread the pre-activations (see tests: _relu_net)
                with torch.no_grad():
                    for g, layer in zip(gains, list(m.net) + [m.last_layer]):
                        layer.linear.weight.mul_(g)
            p = trunc16(m.flat_params()).to(dev)
            call = lambda: F.decode_image(scene.msb, p, K_, D, bc, NL, flags=flags, relu=relu, return_tensor=True,
                                          base_max=scene.msb_max)
            call()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                call()
            a1.record()
            torch.cuda.synchronize()
            ms = a0.elapsed_time(a1) / reps
            flop_px = 2 * (dim_in * bc + (NL - 1) * bc * bc + bc * C_)
            tf = SIDE * SIDE * flop_px / (ms * 1e-3) / 1e12
            return {"config": label, "Mpix_s": SIDE * SIDE / ms / 1e3, "ms_per_scene": ms, "flop_per_pixel": flop_px,
                    "tflops": tf, "tensor_roofline_frac": tf / pk["bf16_burst"]}

        if world == 1:
            other += [
                timed_decode(3, 256, F.Flags(), "configs[2]: D=3 bc256 nl2 (wide tcgen05 kernel, streamed operands)"),
                timed_decode(2, 64, F.Flags(use_coordinates=True, embedding=True, use_colors=False),
                             "configs[3]: USE_COORDINATES+EMBEDDING, USE_COLORS off (dim_in 50)"),
                timed_decode(2, 64, F.Flags(use_coordinates=True, embedding=True, use_colors=True),
                             "configs[3]: USE_COORDINATES+EMBEDDING with colours (dim_in 150)"),
                timed_decode(2, 64, F.Flags(), "north_star ReLU variant (encode.py:75 commented alternative): K=5 D=2 bc64 nl2, "
                                               "ReLU hidden activation on the tensor path", relu=True, gains=(1000.0, 10.0, 40.0)),
            ]
            if not args.no_encode:
                # config 3 encode: D=3 bc256, two epochs timed (training kernel + tensor-core evaluation), scaled to ten
                torch.manual_seed(19920517)
                fl3 = F.Flags()
                tr3 = F.FusedTrainer(LBDRNModel(fl3.dim_in(C_, 3), 256, C_, NL), scene, 3, 1e-3, 8192, 2, flags=fl3, sampler="device")
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r3 = tr3.run()
                torch.cuda.synchronize()
                s3 = time.perf_counter() - t0
                tr3.close()
                other.append({"config": "configs[2] ENCODE: D=3 bc256 nl2, bs 8192, 2 epochs timed (incl. per-epoch evaluation)",
                              "s_per_2_epochs": s3, "s_per_scene_10_epochs_extrapolated": s3 * 5,
                              "us_per_step_incl_eval": s3 / len(r3["losses"]) * 1e6, "final_val_mse": r3["val_mse"][-1]})
                # config 4 encode (README.md:57: USE_COORDINATES + EMBEDDING, USE_COLORS off: dim_in 50), two epochs timed
                torch.manual_seed(19920517)
                fl4 = F.Flags(use_coordinates=True, embedding=True, use_colors=False)
                tr4 = F.FusedTrainer(LBDRNModel(fl4.dim_in(C_, D_), BC, C_, NL), scene, D_, 1e-3, 8192, 2, flags=fl4, sampler="device")
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r4 = tr4.run()
                torch.cuda.synchronize()
                s4 = time.perf_counter() - t0
                tr4.close()
                other.append({"config": "configs[3] ENCODE: USE_COORDINATES+EMBEDDING, USE_COLORS off (dim_in 50), bc64 nl2, bs 8192, "
                                        "2 epochs timed (incl. per-epoch evaluation)",
                              "s_per_2_epochs": s4, "s_per_scene_10_epochs_extrapolated": s4 * 5,
                              "us_per_step_incl_eval": s4 / len(r4["losses"]) * 1e6, "final_val_mse": r4["val_mse"][-1]})

        # configs[4]: 8-band 16-bit 16384 x 16384 GF-6-shaped scene, decode sharded as row stripes over the N ranks (strong
        # scaling: the scene is fixed, every rank takes 16384/N rows + D-row halos), and one configs[1] scene split the same way
        def striped_decode(C, H, W, bits, label, reps):
            rows = LD.stripe_bounds(H, world, rank)
            n_rows = rows[1] - rows[0]
            sc = F.DeviceScene.from_image(make_scene_torch(C, n_rows, W, bits, seed=777 + rank, device=dev), K_)
            torch.manual_seed(19920517)
            p = trunc16(LBDRNModel(fl.dim_in(C, D_), BC, C, NL).flat_params()).to(dev)
            if world > 1:
                dec = LD.StreamedStripeDecoder(H, W, C, D_, sc.msb.dtype, K_, BC, NL, p, fl, dev, slots=1, sub_rows=n_rows)
                dec.preload(sc.msb)
                step = lambda: cur.wait_event(dec.submit())
            else:
                o = torch.empty((C, H, W), dtype=torch.uint16, device=dev)
                dd = sc.desc(D_, BC, NL, fl)
                step = lambda: cabi.check(lib.lbdrn_decode(ctypes.byref(dd), cabi.ptr(sc.msb), cabi.ptr(p), None, cabi.ptr(o),
                                                           cabi.stream_ptr()))
            ms = _timed(step, reps, dev, world)
            flop_px = 2 * (fl.dim_in(C, D_) * BC + (NL - 1) * BC * BC + BC * C)
            tf = H * W * flop_px / (ms * 1e-3) / 1e12
            return {"config": label, "n_gpus": world, "rows_per_gpu": n_rows, "msb_dtype": str(sc.msb.dtype).replace("torch.", ""),
                    "Mpix_s": H * W / ms / 1e3, "ms_per_scene": ms, "flop_per_pixel": flop_px, "tflops_all_gpus": tf,
                    "tensor_roofline_frac_per_gpu": tf / world / pk["bf16_burst"], "scaling": "strong",
                    "per_scene_collectives": "1-element all-reduce(MAX) + one D-row halo swap per seam" if world > 1 else "none"}

        other.append(striped_decode(8, 16384, 16384, 16, "configs[4]: 8-band 16-bit 16384x16384 scene, K=5 D=2 bc64 nl2, "
                                                          f"row stripes over {world} GPU(s)", 3))
        if world > 1:
            other.append(striped_decode(C_, SIDE, SIDE, BITS, f"configs[1] scene (4x8192x8192, 12-bit) split over {world} GPUs "
                                                              "(strong scaling of ONE scene)", 5))

    # ---- CPU baseline (rank 0, N=1 only): oracle port on the host cores, bounded sample ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        side = 3072                                # ~10-20 s of host work on a 16-core box: a bounded sample, not the scene
        t = cpu_decode_sample(side, threads)
        cpu = {"value": side * side / t / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"one decode of a {C_}x{side}x{side} synthetic scene ({t:.1f} s) by oracle/lbdrn_oracle.py",
               "verbatim_reference": cpu_reference_note()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if not has_tc else "f16x2-split/f32-accum",
                "data": "synthetic",
                "config": {"workload": f"configs[1]: synthetic {C_}-band {BITS}-bit {SIDE}x{SIDE} scene per GPU "
                                       f"(row stripes of a {H_total}x{SIDE} scene), K={K_} D={D_} bc{BC} nl{NL}, "
                                       "Sine(w0=30) hidden / Sigmoid head, weights = seeded reference init after fpzip prec16",
                           "l2": "inputs+outputs per step = 805 MB > 126 MB L2 (no flush needed)",
                           "parallelism": f"stripe-sharded decode x{world}"},
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "encode": encode, "other_configs": other,
                "stripe_check": stripe_check}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-encode", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the explanatory lines for the other BASELINE configs")
    ap.add_argument("--encode-epochs", type=int, default=10)
    ap.add_argument("--sampler", choices=["reference", "device"], default="device")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
