"""GPU: the wide (bc 128 / 256) tcgen05 decode kernel of lbdrn_tcw.cu -- BASELINE.json config 3 (D=3 bc256 nl2) --
against the oracle on the same weights, through the C ABI.  Bar as in test_gpu_decode.py: >= 99.99 % of sub-pixels
identical, max |diff| <= 1."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_case, split_stream
import fpzip  # shim
import lbdrn_cabi as cabi
import lbdrn_fused as F
import lbdrn_oracle as O

pytestmark = pytest.mark.gpu


def _check(out, ref, what):
    diff = np.abs(out.astype(np.int64) - ref.astype(np.int64))
    n_bad = int((diff != 0).sum())
    assert diff.max() <= 1, f"{what}: max |diff| = {diff.max()}, {n_bad}/{diff.size} differ"
    assert n_bad <= max(1, int(1e-4 * diff.size)), f"{what}: {n_bad}/{diff.size} sub-pixels differ"


def _stream_params(case):
    meta, img, blob, recon = load_case(case)
    _, tiles = split_stream(blob)
    return meta, np.asarray(fpzip.decompress(tiles[0][0])[0][0][0], dtype=np.float32)


def test_selftest_gemm_n256_and_n16_descriptors():
    """The operand shapes the wide kernel adds: B with 256 rows (LBO 4096 B) and with 16 rows (output layer)."""
    lib = cabi.load()
    rng = np.random.default_rng(5)
    for (N, K) in [(256, 16), (256, 208), (256, 256), (16, 256), (128, 128)]:
        A = rng.integers(-9, 10, size=(128, K)).astype(np.float16)
        B = rng.integers(-9, 10, size=(N, K)).astype(np.float16)
        a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        d = torch.full((128, N), float("nan"), device="cuda")
        cabi.check(lib.lbdrn_selftest_tc_gemm2(cabi.ptr(a), cabi.ptr(b), cabi.ptr(d), N, K, 0, 0, cabi.stream_ptr()))
        torch.cuda.synchronize()
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        assert np.array_equal(d.cpu().numpy().astype(np.float64), ref), (N, K)


@pytest.mark.parametrize("case,shape", [("d3_bc256", (4, 64, 64)), ("d3_bc256", (4, 203, 317)), ("d3_bc256", (4, 8, 2100)),
                                        ("k9_bc128", (4, 150, 131))])
def test_wide_tensor_decode_vs_oracle(case, shape):
    """Config 3 weights (a real reference encode, fpzip prec 16) on scenes with ragged sizes: many tiles per CTA, partial
    tiles on both borders, reflection on all four sides."""
    from synth_scene import make_scene
    meta, params = _stream_params(case)
    C, H, W = shape
    K, D, bc, nl = meta["K"], meta["D"], meta["bc"], meta["nl"]
    img = make_scene(C, H, W, 12, seed=H + 3 * W)
    msb, _ = O.split_msb_lsb(img, K)
    lib = cabi.load()
    d = cabi.make_desc(C, H, W, K, D, bc, nl, F.Flags().bits(), int(msb.max()), msb.max() > 255)
    assert lib.lbdrn_has_tensor_path(ctypes.byref(d)) == 1
    ref = O.decode_image(msb, O.unflatten_params(params, C * (2 * D + 1) ** 2, bc, C, nl), K, D)
    for path in ("tensor", "tensor_fastsin", "auto", "precise"):
        out = F.decode_image(msb, params, K, D, bc, nl, flags=F.Flags(), path=path)
        _check(out, ref, f"{case}/{shape}/{path}")


def test_wide_generic_shape_8_bands_nl3():
    """Table-driven feature offsets (C=8, D=1) and three hidden layers through the streamed-operand ring."""
    from synth_scene import make_scene
    from LBDRNmodel import LBDRNModel
    C, H, W, K, D, bc, nl = 8, 70, 90, 6, 1, 128, 3
    img = make_scene(C, H, W, 12, seed=21)
    msb, _ = O.split_msb_lsb(img, K)
    torch.manual_seed(5)
    flat = LBDRNModel(C * 9, bc, C, nl).flat_params()
    flat = (flat.view(torch.int32) & -65536).view(torch.float32).numpy()          # fpzip prec-16 value map
    ref = O.decode_image(msb, O.unflatten_params(flat, C * 9, bc, C, nl), K, D)
    for path in ("tensor", "tensor_fastsin"):
        _check(F.decode_image(msb, flat, K, D, bc, nl, flags=F.Flags(), path=path), ref, f"c8d1nl3/{path}")


def test_wide_falls_back_to_fp32_kernel_for_inexact_weights():
    """Weights with full fp32 mantissas (-prec 32 streams) are detected on the device: the wide kernel exits and the fp32
    kernel queued behind it decodes the scene -- same parity bar."""
    from synth_scene import make_scene
    meta, params = _stream_params("d3_bc256")
    rng = np.random.default_rng(3)
    flat = (params.astype(np.float64) * (1.0 + 1e-4 * rng.standard_normal(params.size))).astype(np.float32)
    img = make_scene(4, 96, 120, 12, seed=9)
    msb, _ = O.split_msb_lsb(img, 5)
    ref = O.decode_image(msb, O.unflatten_params(flat, 196, 256, 4, 2), 5, 3)
    _check(F.decode_image(msb, flat, 5, 3, 256, 2, flags=F.Flags(), path="tensor"), ref, "inexact/tensor")


@pytest.mark.parametrize("bc,D,shape", [(256, 3, (4, 203, 317)), (256, 3, (4, 1024, 1024)), (256, 2, (4, 300, 500)),
                                        (128, 2, (4, 150, 131)), (128, 3, (4, 257, 513))])
def test_wide_tensor_evaluation_matches_the_fp32_kernel(bc, D, shape):
    """lbdrn_eval_sse at bc 128 / 256 (the per-epoch full-scene MSE of config-3 training, encode.py:105-108) on the wide
    tcgen05 kernel with fp32 weights -- not fp16-exact, so every product also takes the low-order weight operand -- against
    the fp32 kernel (path "precise") on the same weights; and deterministic (three runs, identical bits)."""
    from synth_scene import make_scene
    from LBDRNmodel import LBDRNModel
    C, H, W = shape
    img = make_scene(C, H, W, 12, seed=bc + D + W)
    scene = F.DeviceScene.from_image(img, 5)
    torch.manual_seed(bc + D)
    model = LBDRNModel(C * (2 * D + 1) ** 2, bc, C, 2)
    flat = model.flat_params().cuda()                       # full fp32 weights: low mantissa bits set
    want = F.eval_mse(scene, flat, D, bc, 2, flags=F.Flags(), path="precise")
    got = [F.eval_mse(scene, flat, D, bc, 2, flags=F.Flags()) for _ in range(3)]
    assert got[0] == got[1] == got[2]
    assert abs(got[0] - want) <= 2e-5 * want, (got[0], want)
    # and against the oracle's forward on a crop-sized scene (float64 accumulation of the same squared errors)
    if H * W <= 203 * 317:
        msb, lsb = O.split_msb_lsb(img, 5)
        params = O.unflatten_params(model.flat_params().numpy(), C * (2 * D + 1) ** 2, bc, C, 2)
        X = torch.from_numpy(O.features(msb, D, O.Flags()))
        y = O.forward(params, X).numpy().astype(np.float64)
        ref = float(((y - O.labels(lsb).astype(np.float64)) ** 2).mean())
        assert abs(got[0] - ref) <= 2e-5 * ref, (got[0], ref)


@pytest.mark.parametrize("bc,D,shape", [(256, 3, (4, 203, 317)), (128, 2, (4, 150, 131))])
def test_wide_tensor_decode_of_full_precision_weights(bc, D, shape):
    """A stream whose weights are NOT fp16-exact after scaling (what `-prec 32` produces; encode.py:129) at bc 128 / 256: the
    exact-weights launch exits on the device and its sibling with hi + lo weight operands decodes the scene on the tensor
    cores (no fp32 kernel behind it any more); bar as for every decode test, against the oracle on the same weights."""
    from synth_scene import make_scene
    from LBDRNmodel import LBDRNModel
    C, H, W = shape
    K = 5
    img = make_scene(C, H, W, 12, seed=bc + 7 * D)
    msb, _ = O.split_msb_lsb(img, K)
    torch.manual_seed(bc + D)
    model = LBDRNModel(C * (2 * D + 1) ** 2, bc, C, 2)
    flat = model.flat_params().numpy()                      # full fp32 mantissas
    lib = cabi.load()
    before = lib.lbdrn_launch_count()
    out = F.decode_image(msb, flat, K, D, bc, 2, flags=F.Flags(), path="auto")
    ref = O.decode_image(msb, O.unflatten_params(flat, C * (2 * D + 1) ** 2, bc, C, 2), K, D)
    _check(out, ref, f"bc{bc} D{D} full-precision weights")
    assert lib.lbdrn_launch_count() - before == 4           # two operand preparations + the two sibling launches, nothing else
