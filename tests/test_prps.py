"""PRPS construction: host bookkeeping on CPU, the segmented-mean kernel K0 against the oracle on GPU."""
import numpy as np
import pytest

import ruviii_b200 as rv
from oracle import prps_oracle as PO
from oracle import ruviii_oracle as O
from ruviii_b200 import _cabi, api, prps


def _design(m1=240, n=515, seed=3):
    rng = np.random.default_rng(seed)
    bio = np.array(["basal", "lum_A", "lum_B", "her2"])[rng.integers(0, 4, m1)]
    batch = np.array(["plate_%d" % i for i in rng.integers(0, 5, m1)])
    bio[:7] = "rare"                                                   # a biology that reaches 3 samples in one batch only
    batch[:7] = ["plate_0", "plate_0", "plate_0", "plate_1", "plate_1", "plate_2", "plate_3"]
    libsize = rng.gamma(5.0, 1.0, m1)
    counts = np.round(rng.gamma(2.0, 30.0, (n, m1)))                   # genes x samples
    se = rv.SummarizedExperiment(assays={"counts": counts}, col_names=["s%d" % i for i in range(m1)],
                                 col_data={"subtype": list(bio), "plate": list(batch), "libsize": list(libsize)})
    return se, bio, batch, libsize, counts


def test_categorical_sets_follow_the_table_logic():
    se, bio, batch, _, counts = _design()
    sets, names = prps.prps_sets_categorical(bio, batch, "plate", 3)
    ref, ref_names = PO.prps_categorical(counts, bio, batch, "plate", 3, apply_log=False)
    assert names == ref_names and len(sets) == ref.shape[1]
    assert not any(nm.startswith("rare") for nm in names)              # bio.dist > 1 (prpsForCategoricalUV.R:104-105)
    assert all("_" not in nm.split("||")[0] for nm in names)           # gsub('_', '-')
    for j, idx in enumerate(sets):                                     # the index sets reproduce the oracle's means
        assert np.allclose(counts[:, idx].mean(axis=1), ref[:, j], rtol=1e-14, atol=0)
    with pytest.raises(ValueError, match="not enough samples"):
        prps.prps_sets_categorical(bio, batch, "plate", 1000)


def test_continuous_sets_take_both_tails():
    se, bio, batch, libsize, counts = _design()
    sets, names = prps.prps_sets_continuous(libsize, bio, batch, 3, 10)
    ref, ref_names = PO.prps_continuous(counts, libsize, bio, batch, 3, 10, apply_log=False)
    assert names == ref_names and len(sets) == ref.shape[1] and len(sets) % 2 == 0
    for j, idx in enumerate(sets):
        assert np.allclose(counts[:, idx].mean(axis=1), ref[:, j], rtol=1e-14, atol=0)
    for lo, hi in zip(sets[0::2], sets[1::2]):
        assert libsize[lo].max() <= libsize[hi].min()


def test_argument_errors_match_the_reference():
    se, *_ = _design()
    with pytest.raises(ValueError, match="single categorical variable"):
        rv.prpsForCategoricalUV(se, "counts", ["plate", "subtype"], "subtype", verbose=False)
    with pytest.raises(ValueError, match="categorical source"):
        rv.prpsForCategoricalUV(se, "counts", "libsize", "subtype", verbose=False)
    with pytest.raises(ValueError, match="k cannot be 0"):
        rv.normalise(se, "counts", bio_variable_prps="subtype", uv_variables_prps=["plate"], k=None, verbose=False)


@pytest.mark.gpu
def test_prps_kernel_matches_oracle(engine):
    se, bio, batch, libsize, counts = _design()
    api.set_engine(engine)
    try:
        cat = rv.prpsForCategoricalUV(se, "counts", "plate", "subtype", min_sample_for_prps=3, apply_log=True, pseudo_count=1,
                                      save_se_obj=False, verbose=False)
        ref, names = PO.prps_categorical(counts, bio, batch, "plate", 3, True, 1.0)
        assert cat.row_names == names and O.rel_frob(cat.values, ref.T) < 1e-14
        con = rv.prpsForContinuousUV(se, "counts", True, 1, "libsize", "subtype", "plate", 3, 10, save_se_obj=False, verbose=False)
        ref2, names2 = PO.prps_continuous(counts, libsize, bio, batch, 3, 10, True, 1.0)
        assert con.row_names == names2 and O.rel_frob(con.values, ref2.T) < 1e-14
        # supervisedPRPS + normalise.R:143 stacking
        rv.supervisedPRPS(se, "counts", "subtype", ["plate", "libsize"], "plate", 3, 10, True, 1, verbose=False)
        rd = rv.replicate_data_from_metadata(se)
        assert rd.values.shape[0] == len(names) + len(names2) and rd.row_names == names + names2
        assert list(se.metadata["PRPS"]["supervised"]) == ["bio:subtype||uv:plate||data:counts", "bio:subtype||uv:libsize||data:counts"]
    finally:
        api.set_engine(None)


@pytest.mark.gpu
def test_fused_prps_rows_equal_uploaded_prps_rows(engine):
    """ruviii_set_prps_sets + PS = NULL: replicate.data built on the device from the resident assay gives the same
    RUV-III result as uploading the PRPS rows (log transform first, as prpsForCategoricalUV.R:91 then ruvIII.R:89 do)."""
    se, bio, batch, libsize, counts = _design(m1=300, n=1203, seed=8)
    sets, names = prps.prps_sets_categorical(bio, batch, "plate", 3)
    ref_ps, _ = PO.prps_categorical(counts, bio, batch, "plate", 3, True, 1.0)
    groups = api.group_codes(list(se.col_names) + names)
    ctl = np.zeros(1203, dtype=bool)
    ctl[::7] = True
    prm = _cabi.make_params(apply_log=True, pseudo_count=1.0, want_ps_rows=True)
    a = engine.normalise(prm, counts.T, ref_ps.T, groups, ctl, [2, 3], want_fullalpha=True, want_ps=True)
    engine.set_prps_sets(sets)
    try:
        b = engine.normalise(prm, counts.T, None, groups, ctl, [2, 3], want_fullalpha=True, want_ps=True, ps_on_device=len(sets))
    finally:
        engine.set_prps_sets(None)
    assert O.rel_frob(b["newY"][1], a["newY"][1]) < 1e-12 and O.rel_frob(b["newY_ps"][1], a["newY_ps"][1]) < 1e-12
    ref = O.ruvIII_literal(counts, se.col_names, ref_ps.T, names, ctl, 3, apply_log=True, pseudo_count=1.0)
    assert O.rel_frob(np.vstack([b["newY"][1], b["newY_ps"][1]]), ref["newY"]) < 1e-9
    with pytest.raises(rv.RuviiiError):                                # PS = NULL without registered sets
        engine.normalise(prm, counts.T, None, groups, ctl, [2], ps_on_device=len(sets))


@pytest.mark.gpu
def test_normalise_end_to_end(engine):
    """normalise.R:61-167 with metadata$NCG supplied: PRPS on the device, then ruvIIIMultipleK."""
    se, bio, batch, libsize, counts = _design(m1=300, n=1203, seed=9)
    ncg = np.zeros(1203, dtype=bool)
    ncg[::5] = True
    se.metadata["NCG"] = ncg
    api.set_engine(engine)
    try:
        out = rv.normalise(se, "counts", apply_log=True, pseudo_count=1, bio_variable_prps="subtype", uv_variables_prps=["plate"],
                           batch_variable_prps="plate", k=[1, 3], verbose=False)
    finally:
        api.set_engine(None)
    ps, names = PO.prps_categorical(counts, bio, batch, "plate", 3, True, 1.0)
    for k in (1, 3):
        ref = O.ruvIII_literal(counts, se.col_names, ps.T, names, ncg, k, apply_log=True, pseudo_count=1.0)
        assert O.rel_frob(out.assays["RUV_K%d_on_counts" % k], ref["newY"][:300].T) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("corr_method_vars", [(["plate"], ["subtype"]), (["plate", "libsize"], ["subtype"])])
def test_normalise_resident_matches_stepwise(engine, corr_method_vars, monkeypatch):
    """normalise.R:61-167 as ONE device-resident pipeline (PRPS sets -> K0, control-gene statistics on the resident assay,
    re-plan with the selected ctl, RUV-III) against the step-by-step route that uploads the assay for every step: the same
    metadata$PRPS blocks, the same metadata$NCG, the same RUV_K* assays."""
    uv_ncg, bio_ncg = corr_method_vars
    outs = []
    for resident in ("1", "0"):
        monkeypatch.setenv("RUVIII_NORMALISE_RESIDENT", resident)
        se, bio, batch, libsize, counts = _design(m1=300, n=1203, seed=9)
        api.set_engine(engine)
        try:
            out = rv.normalise(se, "counts", apply_log=True, pseudo_count=1, bio_variable_prps="subtype", uv_variables_prps=["plate"],
                               batch_variable_prps="plate", norm_assay_name_ncg="counts", bio_variables_ncg=bio_ncg,
                               uv_variables_ncg=uv_ncg, no_ncg=200, k=[2, 4], verbose=False)
        finally:
            api.set_engine(None)
        outs.append(out)
    a, b = outs
    assert np.array_equal(a.metadata["NCG"], b.metadata["NCG"]) and a.metadata["NCG"].sum() >= 200
    assert list(a.metadata["PRPS"]["supervised"]) == list(b.metadata["PRPS"]["supervised"])
    for key in a.metadata["PRPS"]["supervised"]:
        pa, pb = a.metadata["PRPS"]["supervised"][key], b.metadata["PRPS"]["supervised"][key]
        assert pa.row_names == pb.row_names and O.rel_frob(pa.values, pb.values) < 1e-13
    for k in (2, 4):
        assert O.rel_frob(a.assays["RUV_K%d_on_counts" % k], b.assays["RUV_K%d_on_counts" % k]) < 1e-11
    ps, names = PO.prps_categorical(counts, bio, batch, "plate", 3, True, 1.0)
    ref = O.ruvIII_literal(counts, se.col_names, ps.T, names, a.metadata["NCG"], 4, apply_log=True, pseudo_count=1.0)
    assert O.rel_frob(a.assays["RUV_K4_on_counts"], ref["newY"][:300].T) < 1e-9
