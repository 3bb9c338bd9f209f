"""Front-end parity through the C ABI: free-running sequences against the oracle (north_star bars: feature IDs
bit-exact, tracked UVs within 0.05 px, status flags equal on >= 99.5 % of features)."""
import numpy as np
import pytest

from oracle import frontend as ofe

pytestmark = pytest.mark.gpu


def _run(fe, synth, n_frames, kw, seed=1000, width=1280, height=560, line_heavy=False, moving_mask=False, lookahead=0):
    seq = synth.SynthSequence(seed=seed, width=width, height=height, n_frames=n_frames, line_heavy=line_heavy,
                              moving_mask=moving_mask)
    oracle = ofe.FrontEnd(ofe.FeConfig(K=seq.K, D=seq.D, **kw))
    gpu = fe.FrontEnd(fe.default_config(width=width, height=height, K=seq.K, D=seq.D, lookahead=lookahead, **kw))
    stats = dict(frames=0, id_equal_frames=0, n_feat=0, n_status_agree=0, max_duv=0.0, line_rows_equal=0, line_frames=0,
                 first_divergence=None)
    for t in range(n_frames):
        img, mask, vps = seq.frame(t), (seq.mask(t) if moving_mask else None), seq.vanishing_points(t)
        prow_o, lrow_o = oracle.feed(seq.timestamp(t), img, mask, vps)
        info = gpu.feed_new_camera(seq.timestamp(t), img, mask, vps)
        rows = gpu.point_rows()
        tr = oracle.klt.trace
        stats["frames"] += 1
        ids_o = np.array([r.id for r in prow_o], np.uint64)
        same_ids = len(ids_o) == len(rows) and np.array_equal(rows["id"], ids_o)
        stats["id_equal_frames"] += int(same_ids)
        if not same_ids and stats["first_divergence"] is None:
            stats["first_divergence"] = t
        if "mask_klt" in tr and same_ids:
            lk = gpu.tap(fe.TAP_LK_LAST, np.float32).reshape(-1, 6)
            assert len(lk) == len(tr["mask_klt"])
            st_o = tr["mask_klt"].astype(bool) & (tr["mask_rsc"].astype(bool) if len(tr["mask_rsc"]) else False)
            st_g = (lk[:, 4] > 0) & (lk[:, 5] > 0)
            stats["n_feat"] += len(lk)
            stats["n_status_agree"] += int((st_o == st_g).sum())
        if same_ids and len(rows):
            uv_o = np.array([[r.u, r.v] for r in prow_o], np.float32)
            stats["max_duv"] = max(stats["max_duv"], float(np.abs(np.stack([rows["u"], rows["v"]], 1) - uv_o).max()))
            un_o = np.array([[r.un, r.vn] for r in prow_o], np.float32)
            stats["max_dun"] = max(stats.get("max_dun", 0.0), float(np.abs(np.stack([rows["un"], rows["vn"]], 1) - un_o).max()))
            d = np.abs(np.stack([rows["u"], rows["v"]], 1) - uv_o).max(1)
            stats.setdefault("duv_all", []).extend(d.tolist())
        lrows, lpts = gpu.line_rows()
        stats["line_frames"] += 1
        if len(lrows) == len(lrow_o) and all(int(a["id"]) == b.id and int(a["D"]) == b.D for a, b in zip(lrows, lrow_o)):
            stats["line_rows_equal"] += 1
        if not same_ids:
            break  # free-running comparison is meaningless after the first ID divergence
    gpu.close()
    d = np.array(stats.pop("duv_all", [0.0]))
    stats["duv_p99"] = float(np.percentile(d, 99))
    stats["duv_gt_0.01"] = int((d > 0.01).sum())
    stats["duv_n"] = int(len(d))
    print(stats)
    return stats


def test_config1_shape_sequence(fe, synth):
    kw = dict(num_features=200, fast_threshold=20, grid_x=5, grid_y=5, min_px_dist=10, pyr_levels=3, win_size=15)
    s = _run(fe, synth, 40, kw)
    assert s["first_divergence"] is None, s
    assert s["n_status_agree"] >= 0.995 * s["n_feat"], s
    assert s["max_duv"] < 0.05, s
    assert s["line_rows_equal"] == s["line_frames"], s


def test_config2_shape_sequence(fe, synth):
    kw = dict(num_features=400, fast_threshold=20, grid_x=5, grid_y=5, min_px_dist=10, pyr_levels=4, win_size=15)
    s = _run(fe, synth, 40, kw, seed=1001)
    assert s["first_divergence"] is None, s
    assert s["n_status_agree"] >= 0.995 * s["n_feat"], s
    assert s["max_duv"] < 0.05, s


def test_moving_mask_sequence(fe, synth):
    kw = dict(num_features=200, fast_threshold=20, grid_x=5, grid_y=3, min_px_dist=10, pyr_levels=5, win_size=15)
    s = _run(fe, synth, 25, kw, seed=1002, moving_mask=True)
    assert s["first_divergence"] is None, s
    assert s["max_duv"] < 0.05, s
