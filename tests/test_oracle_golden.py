"""CPU: the oracle restatement (oracle/lbdrn_oracle.py) against the fixtures minted from the UNMODIFIED reference."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import CODEC_CASES, GOLD, case_flags, load_case, read_base, split_stream
import lbdrn_oracle as O
import fpzip  # shim


def test_header_kat():
    for k in json.load(open(os.path.join(GOLD, "header_kat.json"))):
        a = k["args"]
        b = O.pack_header(a["split_ratio"], a["width"], a["height"], a["K"], a["bc"], a["nl"], a["D"],
                          a["nn_bytes_list"], a["base_bytes_list"])
        assert b.hex() == k["hex"]
        assert list(O.unpack_header(b)) == k["parsed"]


def test_header_appendix_b_vector():
    # SURVEY.md Appendix B.1, produced by the reference's own write_image_header
    b = O.pack_header(1, 2048, 2048, 5, 64, 2, 2, [19300], [1234567])
    assert b.hex() == "0f01080008005262004b640012d687"


def test_seeded_init_kat():
    for k in json.load(open(os.path.join(GOLD, "init_kat.json"))):
        a = k["args"]
        torch.manual_seed(19920517)
        p = O.init_params(a["dim_in"], a["dim_hidden"], a["dim_out"], a["num_layers"])
        flat = O.flatten_params(p)
        assert flat.size == k["n"] == O.n_params(a["dim_in"], a["dim_hidden"], a["dim_out"], a["num_layers"])
        assert hashlib.sha256(flat.tobytes()).hexdigest() == k["sha256"]
        draws = [int(torch.empty((), dtype=torch.int64).random_()) for _ in range(2)]
        assert draws == k["next_draws"]


def test_feature_layout_kat():
    from synth_scene import make_scene
    idx = json.load(open(os.path.join(GOLD, "features_index.json")))
    for name, meta in idx.items():
        g = np.load(os.path.join(GOLD, f"features_{name}.npz"))
        sc = meta["scene"]
        img = make_scene(sc["C"], sc["H"], sc["W"], sc["bits"], seed=sc["seed"])
        msb, lsb = O.split_msb_lsb(img, meta["K"])
        assert np.array_equal(msb, g["base"]) and msb.dtype == g["base"].dtype
        f = O.features(msb, meta["D"], O.Flags(**meta["flags"]))
        assert f.shape == tuple(meta["shape"])
        assert np.array_equal(f, g["features"]), name
        assert np.array_equal(O.labels(lsb), g["labels"]), name
        # row-chunked evaluation gives the same rows
        f2 = np.concatenate([O.features(msb, meta["D"], O.Flags(**meta["flags"]), r, min(sc["H"], r + 4))
                             for r in range(0, sc["H"], 4)])
        assert np.array_equal(f2, f)


def test_forward_kat():
    g = np.load(os.path.join(GOLD, "forward_d100_bc64.npz"))
    p = O.unflatten_params(g["params"], 100, 64, 4, 2)
    with torch.no_grad():
        y = O.forward(p, torch.from_numpy(g["x"]))
    assert np.array_equal(y.numpy(), g["y"])


def test_forward_relu_kat():
    """The ReLU variant (encode.py:75 / decode.py:108, commented alternative): the oracle's forward, loss and autograd
    gradients against the reference's own LBDRNModel(activation=nn.ReLU()) + LBDRNLoss."""
    g = np.load(os.path.join(GOLD, "forward_relu_d100_bc64.npz"))
    p = [t.requires_grad_(True) for t in O.unflatten_params(g["params"], 100, 64, 4, 2)]
    y = O.forward(p, torch.from_numpy(g["x"]), relu=True)
    assert np.array_equal(y.detach().numpy(), g["y"])
    assert 0.2 < float((y.detach() > 0.5).float().mean()) < 0.8          # the fixture exercises both sides of the head
    loss = torch.nn.functional.mse_loss(y, torch.from_numpy(g["t"]))
    loss.backward()
    assert np.float32(loss.item()) == g["loss"]
    assert np.array_equal(np.concatenate([t.grad.numpy().reshape(-1) for t in p]), g["grad"])


@pytest.mark.parametrize("name", CODEC_CASES + ["sr2_tiles"])
def test_decode_matches_reference_bit_exactly(name):
    meta, img, blob, recon = load_case(name)
    hdr, tiles = split_stream(blob)
    n, sr, W, H, K, bc, nl, D = hdr[:8]
    assert (W, H, K, bc, nl, D) == (meta["W"], meta["H"], meta["K"], meta["bc"], meta["nl"], meta["D"])
    flags = case_flags(meta, O.Flags)
    out = np.zeros_like(recon)
    tw, th = W // sr, H // sr
    for t, (nn, base) in enumerate(tiles):
        i, j = divmod(t, sr)
        b = read_base(base)
        params = O.unflatten_params(fpzip.decompress(nn)[0][0][0], flags.dim_in(b.shape[0], D), bc, b.shape[0], nl)
        rec = O.decode_image(b, params, K, D, flags)
        out[:, i * th:i * th + rec.shape[1], j * tw:j * tw + rec.shape[2]] = rec
    assert np.array_equal(out, recon)
    mse, psnr, bpsp = O.quality(img, out, len(blob))
    assert abs(psnr - meta["psnr"]) < 1e-3 and abs(bpsp - meta["bpsp"]) < 1e-9


def test_training_matches_reference_step_for_step():
    meta, img, blob, recon = load_case("k5d2_small")
    torch.manual_seed(19920517)
    msb, lsb = O.split_msb_lsb(img, meta["K"])
    r = O.train(msb, lsb, meta["D"], meta["bc"], meta["nl"], 1e-3, meta["bs"], meta["e"])
    assert len(r["losses"]) == len(meta["losses"])
    assert np.max(np.abs(np.array(r["losses"]) - np.array(meta["losses"]))) < 1e-7
    assert np.allclose(r["mses"], meta["val_mse"], rtol=1e-6)
    assert r["best_epoch"] == meta["best_epoch"][0]
    # the nn sub-stream the reference wrote holds exactly these parameters after the fpzip value map
    _, tiles = split_stream(blob)
    assert np.array_equal(O.fpzip_value_map(O.flatten_params(r["params"]), 16), fpzip.decompress(tiles[0][0])[0][0][0])


def test_lr_schedule():
    assert [O.lr_at_epoch(1e-3, e, 10) for e in (1, 3, 4, 6, 7, 9, 10)] == pytest.approx(
        [1e-3, 1e-3, 1e-4, 1e-4, 1e-5, 1e-5, 1e-6], rel=1e-12)
    assert O.lr_at_epoch(1e-3, 2, 1) == pytest.approx(1e-4)     # step_size = max(1, int(1/3)) = 1


KAT_PERM_10_SEED1 = [9, 8, 0, 1, 6, 7, 3, 5, 2, 4]
KAT_PERM_1000_HEAD = [425, 988, 249, 411, 710, 149, 364, 339]


def test_device_permutation_restatement_is_a_bijection_with_known_answers():
    """The restatement of lbdrn_randperm: a permutation for every n, seed-dependent, and pinned by known answers (so that
    the GPU parity test compares against a fixed target, not against a moving one)."""
    import lbdrn_oracle as O
    for n in (1, 2, 3, 5, 64, 1000, 4099):
        p = O.device_permutation(n, 12345)
        assert np.array_equal(np.sort(p), np.arange(n))
    assert not np.array_equal(O.device_permutation(1000, 1), O.device_permutation(1000, 2))
    assert O.device_permutation(10, 1).tolist() == KAT_PERM_10_SEED1
    assert O.device_permutation(1000, 0xDEADBEEFCAFEF00D)[:8].tolist() == KAT_PERM_1000_HEAD
