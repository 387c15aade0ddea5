"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and exports every
symbol include/mtm_b200.h declares; the Python binding table covers the same set; the product
path refuses to run (loudly) without a CUDA device instead of falling back to anything."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mtm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import mtm_b200

    if not os.path.exists(mtm_b200._capi.LIB_PATH):
        mtm_b200.build()
    lib = ctypes.CDLL(mtm_b200._capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 19
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mtm_b200.h but not exported"
    assert set(names) == set(mtm_b200._capi.SIGNATURES), "binding table and header disagree"
    lib.mtm_abi_version.restype = ctypes.c_int
    assert lib.mtm_abi_version() == 1
    assert lib.mtm_param_stride(4, 4) == 85 and lib.mtm_stats_stride(4, 4) == 106
    assert lib.mtm_param_stride(2, 4) == 1 + 2 + 12 + 8 + 16


def test_layout_helpers_match_oracle_packing():
    from oracle import dense as od

    d, l = 3, 2
    rng = np.random.default_rng(0)
    par = od.Params([0.4, 0.6], [rng.normal(size=d)] * 2, [np.eye(d)] * 2, [rng.normal(size=(d, d))] * 2,
                    [np.eye(d)] * 2, [rng.normal(size=(d, l))] * 2, [np.eye(l)] * 2)
    P = par.packed()
    import mtm_b200

    assert P.shape == (2, mtm_b200._capi.load().mtm_param_stride(d, l))
    back = od.Params.from_packed(P, d, l)
    assert np.array_equal(back.packed(), P)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mtm_b200
    from mtm_b200.framework import marginalizable_mixture_model as mixmodel

    mdl = mixmodel.MMLinGaussSS_marginalizable(2, np.zeros((3, 20, 2)), np.zeros((3, 20, 2)))
    # host-side construction works (initial assignment, hash); every numeric call must raise
    assert mdl.cluster_assignment.shape == (20,)
    for call in (mdl.train, mdl.E_step, mdl.M_step, mdl.mle_cluster_assignment, mdl.e_complete_data_log_lik):
        with pytest.raises(mtm_b200.MtmError):
            call()
    # mtm_create itself reports the missing device
    lib = mtm_b200._capi.load()
    h = ctypes.c_void_p()
    assert lib.mtm_create(0, None, ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.mtm_last_error(None)


def test_host_side_matches_reference_ctor(g_em, g_adni):
    """Initial assignments and the data hash are host code shared with the reference
    (MMM:128-188): same numpy stream, same md5."""
    from mtm_b200.framework import marginalizable_mixture_model as mixmodel

    for key in g_em["cases"]:
        n, T, d, l, n_c, seed, _ = (int(v) for v in g_em[f"{key}_meta"])
        mdl = mixmodel.MMLinGaussSS_marginalizable(n_c, g_em[f"{key}_z"], g_em[f"{key}_x"], random_seed=seed,
                                                   alpha=float(g_em[f"{key}_alpha"]))
        assert np.array_equal(mdl.cluster_assignment, g_em[f"{key}_assign0"])
        assert mdl.n_free_params == n_c * (1 + d + d * d + d * l) + n_c * (d * (d + 1) + l * (l + 1) // 2)
    mdl = mixmodel.MMLinGaussSS_marginalizable(4, g_adni["ztr"], g_adni["xtr"], init="k-means", alpha=1.0)
    assert mdl.hex_hash == str(g_adni["hex_hash"])
    km = mixmodel.MMLinGaussSS_marginalizable(4, g_adni["ztr"], g_adni["xtr"], random_seed=0, init="kmeans", alpha=1.0)
    assert np.array_equal(km.cluster_assignment, g_adni["km_assign0"])


def test_state_space_helpers(g_ssm):
    from mtm_b200.framework import marginalizable_state_space_model as ssm

    for k, (T, d, l) in enumerate(g_ssm["shapes"]):
        g = lambda name: g_ssm[f"s{k}_{name}"]
        assert np.allclose(ssm.mm(T, g("m"), g("A"), g("H")), g("mean"), rtol=1e-12, atol=1e-12)
        assert np.allclose(ssm.CC(T, g("S"), g("A"), g("G"), g("H"), g("L")), g("cov"), rtol=1e-12, atol=1e-12)
        assert np.allclose(ssm.mmX(T, g("m"), g("A"), g("H")), g("meanx"), rtol=1e-12, atol=1e-12)
        assert np.allclose(ssm.CXX(T, g("S"), g("A"), g("G"), g("H"), g("L")), g("covx"), rtol=1e-12, atol=1e-12)


def test_headline_kernel_keeps_uniform_parameter_operands():
    """Regression guard for the E-step code shape: in the (d, l) = (4, 4) kernel the
    model parameters must reach the DFMAs as uniform-register operands.  tools/sass_reuse.py models the
    vector-register reads per FP64 instruction of the time loop from the SASS: ~0.89 in the intended form,
    ~0.79 when ptxas falls back to vector-register parameter loads."""
    import shutil
    import subprocess
    import sys

    import mtm_b200

    obj = os.path.join(ROOT, "multimodal-trajectory-modeling_b200", "csrc", "_build", "estep.o")
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    if not os.path.exists(obj):
        mtm_b200.build()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_reuse.py"), obj, "ILi4ELi4ELb1ELb1E"],
                         capture_output=True, text=True, check=True).stdout
    eff = float(re.search(r"pipe efficiency ([0-9.]+)", out).group(1))
    assert eff >= 0.85, out
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert re.search(r"DFMA R\d+, R\d+(\.reuse)?, UR\d+, R", sass), "no uniform-register DFMA operands"
    # both (4, 4) kernels (joint and observation-only): the parameter bank c[0x3] is read through the uniform
    # datapath (LDCU.64 UR, c[0x3][UR+imm]); the fallback form reads it with LDC.64 R, c[0x3][R+imm]
    # ... and the delta-statistics twin of the joint kernel (estep_delta_kernel<4, 4, CONSTP>)
    for kern in ("estep_kernelILi4ELi4ELb1ELb1E", "estep_kernelILi4ELi4ELb0ELb1E", "estep_delta_kernelILi4ELi4ELb1E"):
        body = next(b for b in re.split(r"\n\s*Function : ", sass) if kern in b.split("\n")[0])
        n_u = len(re.findall(r"LDCU\.64 UR\d+, c\[0x3\]", body))
        n_v = len(re.findall(r"LDC\.64 R\d+, c\[0x3\]", body))
        assert n_u >= 80 and n_u > 2 * n_v, (kern, n_u, n_v)


def test_import_as_reference_package_name():
    """with the package directory first on sys.path the reference's own import line
    `from framework import marginalizable_mixture_model as mixmodel` resolves to the drop-in class."""
    import subprocess
    import sys

    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); "
        "from framework import marginalizable_mixture_model as mixmodel; "
        "m = mixmodel.MMLinGaussSS_marginalizable(n_clusters=2, states=np.zeros((3, 12, 2)), observations=np.zeros((3, 12, 1))); "
        "print(type(m).__module__, m.cluster_assignment.shape, len(m.hex_hash))"
    ) % os.path.join(ROOT, "multimodal-trajectory-modeling_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr[-800:]
    assert out.stdout.split()[0] == "framework.marginalizable_mixture_model" and "(12,)" in out.stdout


def test_extended_host_helpers():
    """framework_extended host-side pieces that need no GPU: take_finite_along_axis (the example of
    util_state_space.py:353-363) and the constructor's bookkeeping of StateSpaceMixtureModel."""
    import mtm_b200  # noqa: F401
    from mtm_b200.framework_extended import state_space_model_linear_gaussian as xl
    from mtm_b200.framework_extended import state_space_model_mixture as xm

    eg = np.concatenate([np.arange(7), np.repeat(np.nan, 3), np.arange(2)]).reshape(6, 2)
    assert np.array_equal(xm.take_finite_along_axis(eg, axis=0), np.array([[0.0, 1.0], [2.0, 3.0], [4.0, 5.0], [0.0, 1.0]]))
    rng = np.random.default_rng(0)
    z, x = rng.normal(size=(3, 20, 2)), rng.normal(size=(3, 20, 1))
    m = xm.StateSpaceMixtureModel(n_clusters=3, data=(z, x), component_model=xl.StateSpaceLinearGaussian,
                                  rng=np.random.default_rng(7))
    assert (m.n_timesteps, m.n_data, m.d_states, m.d_observations) == (3, 20, 2, 1)
    assert np.array_equal(m.cluster_assignment, np.random.default_rng(7).integers(3, size=20))
    assert m.correspondence == {0: "A", 1: "B", 2: "C"} and m.inverse_correspondence["B"] == 1
    assert np.allclose(m.cluster_propensities, 1 / 3) and len(m.cluster_models) == 3 and len(m.hex_hash) == 32
