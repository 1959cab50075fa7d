"""GPU parity of K4 (fusion + compound rule + argmax) and its alignment glue, through the C ABI.
Contract (BASELINE.json north star): bit-exact given identical input probabilities."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import fusion as of
from oracle.make_golden import FUSION_CONFIGS, fusion_weights, synthetic_fusion_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_k4_bit_exact_vs_reference_golden(cuda_lib, golden):
    from avcer_b200 import ops

    g = golden["fusion"]
    p_vs, p_vd, p_a = _t(g["p_vs"]), _t(g["p_vd"]), _t(g["p_a"])
    for tag, cwt, cm in FUSION_CONFIGS:
        w1, w2 = fusion_weights(tag)
        got = ops.fuse_compound(p_vs, p_vd, p_a, w1, w2, cwt, cm).cpu().numpy()
        assert np.array_equal(got, g[f"labels_{tag}_{int(cwt)}_{int(cm)}"]), (tag, cwt, cm)
    got = ops.fuse_compound(p_vs.double(), p_vd.double(), p_a.double(), None, [1, 1, 1], False, True).cpu().numpy()
    assert np.array_equal(got, g["labels_none_f64_0_1"])


def test_k4_bit_exact_vs_oracle_random_with_specials(cuda_lib):
    from avcer_b200 import ops

    rng = np.random.default_rng(3)
    n = 20011
    ps = [rng.dirichlet(np.ones(7) * a, size=n).astype(np.float32) for a in (0.3, 1.0, 5.0)]
    ps[0][::97] = np.nan                       # NaN rows (audio frames covered only by the empty window)
    ps[1][::89] = 0.0                          # zero VS rows
    ps[2][::83] = np.float32(1 / 7)            # on the mask threshold
    ps[1][5::101] = ps[1][4::101][: len(ps[1][5::101])]
    for tag, cwt, cm in FUSION_CONFIGS:
        w1, w2 = fusion_weights(tag)
        ref = np.stack(of.fuse_labels(ps[0], ps[1], ps[2], w1, w2, cwt, cm))
        got = ops.fuse_compound(_t(ps[0]), _t(ps[1]), _t(ps[2]), w1, w2, cwt, cm).cpu().numpy()
        assert np.array_equal(got, ref), (tag, cwt, cm, int((got != ref).sum()))


def test_k4_full_size_chunk_consistency(cuda_lib):
    """C4 size (1.5 M frames): labels of the whole launch == labels of independent slices, and the
    sampled rows agree bit-exactly with the oracle."""
    from avcer_b200 import ops

    n = 1_500_000
    g = torch.Generator(device=DEV).manual_seed(0)
    ps = [torch.softmax(torch.randn(n, 7, device=DEV, generator=g) * s, 1).contiguous() for s in (1.0, 2.0, 0.5)]
    w1, w2 = fusion_weights("w3")
    full = ops.fuse_compound(ps[0], ps[1], ps[2], w1, w2, False, True)
    for lo, hi in ((0, 1), (123, 70001), (n - 513, n)):
        part = ops.fuse_compound(ps[0][lo:hi].contiguous(), ps[1][lo:hi].contiguous(), ps[2][lo:hi].contiguous(), w1, w2, False, True)
        assert torch.equal(part, full[:, lo:hi])
    idx = torch.randint(0, n, (5000,), device=DEV, generator=g)
    ref = np.stack(of.fuse_labels(*[p[idx].cpu().numpy() for p in ps], w1, w2, False, True))
    assert np.array_equal(full[:, idx].cpu().numpy(), ref)
    assert ops.fuse_compound(ps[0][:0].contiguous(), ps[1][:0].contiguous(), ps[2][:0].contiguous(), w1, w2, False, True).shape == (4, 0)


def test_softmax7_close_to_numpy(cuda_lib):
    from avcer_b200 import ops

    rng = np.random.default_rng(1)
    x = (rng.standard_normal((4099, 8)) * 3).astype(np.float32)
    x[7, :7] = 0.0
    got = ops.softmax7(_t(x)).cpu().numpy()
    ref = of.softmax(x[:, :7])
    assert np.abs(got - ref).max() < 2e-7          # tolerance: expf vs numpy's SIMD exp differ by <= 1-2 ulp
    assert np.array_equal(got[7], np.full(7, np.float32(1) / np.float32(7)))
    x64 = rng.standard_normal((100, 7))
    assert np.abs(ops.softmax7(_t(x64)).cpu().numpy() - of.softmax(x64)).max() < 1e-15


def test_window_to_frame_mean_bit_exact_vs_pandas(cuda_lib, golden):
    from avcer_b200 import ops, pipeline

    g = golden["audio"]
    for ncls in (8, 7):
        for tag in "abc":
            L, fps, pad, step = (int(v) for v in g[f"a{ncls}_{tag}_meta"])
            ap = pipeline.plan_audio(L, fps, step / 1000)
            ids = g[f"a{ncls}_{tag}_frame_ids"]
            n = int(ids.max()) + 1
            out = ops.window_to_frame_mean(_t(g[f"a{ncls}_{tag}_window_logits"]), _t(ap.f_lo.astype(np.int32)),
                                           _t(ap.f_hi.astype(np.int32)), n).cpu().numpy()
            assert np.array_equal(out[ids], g[f"a{ncls}_{tag}_frame_means"], equal_nan=True), (ncls, tag)


def test_run_get_c_expr_db_pred_dropin(cuda_lib, golden):
    """The reference entry point on DataFrames: labels identical to the stock run.get_c_expr_db_pred
    for every weight table / rule combination (softmax differences stay below the decision margins of
    this fixture; K4 itself is covered bit-exactly above)."""
    from avcer_b200 import run as arun

    g = golden["fusion"]
    stat_df, dyn_df, audio_df = synthetic_fusion_inputs()
    for tag, cwt, cm in FUSION_CONFIGS:
        w1, w2 = fusion_weights(tag)
        av, vs, vd, a, loc = arun.get_c_expr_db_pred(stat_df.copy(), dyn_df.copy(), audio_df.copy(), "clip", w1, w2, cwt, cm, False)
        ref = g[f"labels_{tag}_{int(cwt)}_{int(cm)}"]
        agree = np.mean(np.stack([av, vs, vd, a]) == ref)
        assert agree >= 0.999, (tag, cwt, cm, agree)
        assert av.dtype == np.int64 and loc[0] == "clip/00001.jpg" and len(loc) == len(stat_df)
    # the aligned probabilities themselves
    p_vs, p_vd, p_a, _ = arun.aligned_probabilities(stat_df, dyn_df, audio_df, "clip")
    assert np.array_equal(p_vs.cpu().numpy(), g["p_vs"])
    assert np.abs(p_vd.cpu().numpy() - g["p_vd"]).max() < 2e-7
    assert np.nanmax(np.abs(p_a.cpu().numpy() - g["p_a"])) < 2e-7
    assert np.array_equal(np.isnan(p_a.cpu().numpy()), np.isnan(g["p_a"]))


def test_data_utils_dropins(cuda_lib, golden):
    from avcer_b200.data import utils as du

    g = golden["fusion"]
    w1, w2 = fusion_weights("w3")
    fused = g["p_vs"] * w1[0] * w2[0] + g["p_vd"] * w1[1] * w2[1] + g["p_a"] * w1[2] * w2[2]
    com = {"a": [3, 6], "b": [4, 6], "c": [5, 6], "d": [2, 6], "e": [1, 6], "f": [3, 5], "g": [1, 5]}
    dw = {1: 5, 2: 6, 3: 5, 4: 6, 5: 4, 6: 2}
    for cwt, cm in ((True, False), (False, True), (True, True), (False, False)):
        got = du.get_compound_expression(fused, com, dw, cwt, cm)
        assert np.array_equal(got, g[f"scores_w3_{int(cwt)}_{int(cm)}"], equal_nan=True)
    assert du.get_image_location("v", "000012.jpg") == "v/00013.jpg"
    with pytest.raises(ZeroDivisionError):
        du.pad_wav(torch.zeros(0), 10)


def test_weight_search_dropins_match_reference(cuda_lib, golden):
    """the Dirichlet / grid weight searches of data/utils.py:138-209, every candidate
    evaluated in one launch; selected weights identical to the stock functions (golden) and per-candidate
    confusion matrices identical to a numpy evaluation."""
    from avcer_b200 import ops
    from avcer_b200.data import utils as du
    from oracle.make_golden import weight_search_inputs

    g = golden["weight_search"]
    gt, preds = weight_search_inputs()
    preds_l = [p.tolist() for p in preds]
    np.random.seed(42)
    bw = du.get_weights_prob_model(gt.tolist(), preds_l, 60, 7)
    assert np.array_equal(bw, g["prob_best"])
    grid = g["grid"].tolist()
    assert du.get_weights_av_model(grid, gt.tolist(), preds_l) == g["av_best"].tolist()
    assert du.get_weights_v_model(grid, gt.tolist(), preds_l[:2]) == g["v_best"].tolist()
    # confusion counts of 300 random candidates against numpy, incl. a ragged frame count
    rng = np.random.default_rng(5)
    W = rng.dirichlet(np.ones(3), size=(300, 7)).transpose(0, 2, 1).copy()
    P = np.stack(preds)[:, :577]
    cm = ops.weight_search_confusion(_t(P), _t(gt[:577].astype(np.int32)), _t(W)).cpu().numpy()
    for w in (0, 17, 299):
        final = P[0] * W[w, 0]
        final += P[1] * W[w, 1]
        final += P[2] * W[w, 2]
        ref = np.zeros((7, 7), dtype=np.int64)
        np.add.at(ref, (gt[:577], np.argmax(final, axis=-1)), 1)
        assert np.array_equal(cm[w], ref)
    assert cm.sum() == 300 * 577


def test_get_pred_av_dropin_matches_reference(cuda_lib, golden, tmp_path, monkeypatch):
    """Row a18: avcer_b200.get_pred_av.get_c_expr_db_pred on the CSV files (float64 tables: per-frame audio means and
    softmax in float64 like pandas / numpy do) writes exactly the labels the unmodified reference wrote."""
    import pandas as pd

    from avcer_b200 import get_pred_av
    from oracle.make_golden import PRED_AV_CONFIGS, pred_av_tables, pred_av_weights, write_pred_av_files

    g = golden["pred_av"]
    tables, fmt = pred_av_tables()
    root = str(tmp_path / "preds")
    fmt_path, path_preds = write_pred_av_files(root, tables, fmt)
    monkeypatch.chdir(tmp_path)
    for i, (tag, w2, cwt, cm) in enumerate(PRED_AV_CONFIGS):
        w1, w2v = pred_av_weights(tag, w2)
        labels, locs = get_pred_av.get_c_expr_db_pred(fmt_path, root, path_preds, list(tables), w1, w2v, tag, f"cfg{i}", cwt, cm)
        assert len(locs) == int(g["n_locations"]) and np.array_equal(np.asarray(labels), g[f"labels_{i}"]), (tag, w2, cwt, cm)
        txt = pd.read_csv(tmp_path / "src" / "pred_results" / "DF_C_EXPR_DB" / f"C_EXPR_DB_{tag}_sd_cfg{i}_{cwt}_{cm}.txt")
        assert list(txt.iloc[:, 0]) == locs and np.array_equal(txt.iloc[:, 1].to_numpy(), g[f"labels_{i}"])


def test_fused_argmax_matches_numpy_semantics(cuda_lib):
    """get_pred_av.get_metrics' fusion (get_pred_av.py:34-40) on the GPU: bit-exact labels against the reference's own
    numpy expression, including ties (first maximum), zero rows and NaN rows."""
    from avcer_b200 import get_pred_av, get_weights_matrices as gwm

    rng = np.random.default_rng(9)
    n = 5003
    preds = [rng.dirichlet(np.ones(7) * 0.5, size=n) for _ in range(3)]
    preds[0][10] = preds[1][10] = preds[2][10] = 0.0
    preds[1][11, 3] = np.nan
    preds[0][12] = preds[0][12][::-1].copy(); preds[1][12] = preds[0][12]; preds[2][12] = preds[0][12]
    for p in preds:
        p[13] = 1.0 / 7.0
    w1 = np.asarray(gwm.class_weights(gwm.weights_3))
    for w2 in ([1, 1, 1], gwm.model_weights(gwm.weights_3)):
        final = preds[0] * w1[0] * w2[0]
        for i in range(1, 3):
            final += preds[i] * w1[i] * w2[i]
        want = np.argmax(final, axis=-1)
        got = get_pred_av.fused_argmax(preds, w1, w2)
        assert got.dtype == np.int32 and np.array_equal(got, want)
