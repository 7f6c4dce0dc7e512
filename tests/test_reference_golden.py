"""Golden fixtures produced by THE REFERENCE'S OWN CODE (tests/golden/ref/*.npz, written by
tests/golden/make_golden_ref.py from oracle/_ref/ref_driver = the unmodified als_CP.cxx / als_Tucker.cxx / common.cxx
compiled against the CTF stand-in) against (a) the oracle restatement, on CPU, and (b) the CUDA path through the C++
host drivers and the C ABI, on the GPU.  Neither needs /root/reference or oracle/_ref at run time.

Bar (BASELINE.json): residual at every print point within 1e-10 relative to ||V||, factors within 1e-8 (Tucker:
projectors W W^T within 1e-7, the columns carry an arbitrary sign), identical DT<->PP switching iterations."""
import importlib
import os

import numpy as np
import pytest

from oracle import pp_oracle as o

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref")
NAMES = sorted(x[:-4] for x in os.listdir(GOLD) if x.endswith(".npz"))
FIT_RTOL, FACTOR_TOL = 1e-10, 1e-8


def load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    c = dict(op=str(g["op"]), lens=tuple(int(v) for v in g["lens"]), R=int(g["R"]), vnorm=float(g["vnorm"]),
             rows=g["rows"], events=[tuple(int(x) for x in r) for r in g["events"].tolist()], g=g)
    for k, d in (("tol_init", 0.01), ("maxiter", 20), ("resprint", 10), ("lambda_", 0.0), ("ratio_step", 1.0),
                 ("update_pct", 1.0)):
        c[k] = type(d)(g[k]) if k in g else d
    c["W"] = [g["W%d" % i] for i in range(len(c["lens"]))]
    return c


def check_rows(got, c, tucker=False):
    got = np.array([(r[0], r[1], r[2], r[3]) for r in got], dtype=float)
    ref = c["rows"]
    assert got.shape == ref.shape
    assert np.array_equal(got[:, 0], ref[:, 0]) and np.array_equal(got[:, 2], ref[:, 2])
    assert np.allclose(got[:, 3], ref[:, 3], rtol=0, atol=FIT_RTOL * c["vnorm"])
    assert np.allclose(got[:, 1], ref[:, 1], rtol=0 if tucker else 1e-9, atol=1e-9 * c["vnorm"])


def proj_err(A, B):
    return np.abs(A @ A.T - B @ B.T).max()


def test_fixture_set_is_complete():
    ops = {load(n)["op"] for n in NAMES}
    assert ops == {"alsCP_DT", "alsCP_PP", "alsCP_PP_partupdate", "alsTucker_DT", "alsTucker_PP", "hosvd"}
    assert len(NAMES) >= 15


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_fixture(name):
    c = load(name)
    lens, R, N, vnorm = c["lens"], c["R"], len(c["lens"]), c["vnorm"]
    kinds = {"DT": 0, "PP": 1}
    if c["op"] == "hosvd":
        V = o.make_tensor_r2(lens)
        core, W = o.hosvd(V, [int(x) for x in c["g"]["ranks"]])
        for a, b in zip(W, c["W"]):
            assert proj_err(a, b) < 1e-9
        assert abs(np.linalg.norm(core) - float(c["g"]["core_norm"])) < 1e-10 * vnorm
        return
    if c["op"].startswith("alsTucker"):
        V = o.make_tensor_r2(lens)
        assert abs(np.linalg.norm(V) - vnorm) < 1e-12 * vnorm
        W = [c["g"]["W_init%d" % i].copy() for i in range(N)]
        core = o.TTMc(V, W, -1)
        if c["op"] == "alsTucker_DT":
            _, rows, core = o.alsTucker_DT(V, core, W, 1e-10 * vnorm, c["maxiter"], resprint=c["resprint"])
            rows = [(r[0], r[1], 0, r[2]) for r in rows]
            events = []
        else:
            _, rows, events, _, core = o.alsTucker_PP(V, core, W, 1e-10 * vnorm, c["tol_init"], c["maxiter"],
                                                      resprint=c["resprint"])
        assert [(kinds[k], it) for k, it in events] == c["events"]
        check_rows(rows, c, tucker=True)
        for a, b in zip(W, c["W"]):
            assert proj_err(a, b) < 1e-7
        return
    V, _ = o.make_tensor_r(lens, R)
    assert abs(np.linalg.norm(V) - vnorm) < 1e-12 * vnorm
    W, G = o.init_factors(lens, R), o.init_grad(lens, R)
    if c["op"] == "alsCP_DT":
        tr = o.Trace()
        o.alsCP_DT(V, W, G, 1e-10 * vnorm, c["maxiter"], lam=c["lambda_"], resprint=c["resprint"], trace=tr)
    elif c["op"] == "alsCP_PP":
        _, tr = o.alsCP_PP(V, W, G, 1e-10 * vnorm, c["tol_init"], c["maxiter"], lam=c["lambda_"],
                           ratio_step=c["ratio_step"], resprint=c["resprint"])
    else:
        _, tr = o.alsCP_PP_partupdate(V, W, G, 1e-10 * vnorm, c["tol_init"], c["maxiter"], lam=c["lambda_"],
                                      ratio_step=c["ratio_step"], update_percentage=c["update_pct"],
                                      resprint=c["resprint"])
    assert [(kinds[k], it) for k, it in tr.events] == c["events"]
    check_rows(tr.rows, c)
    for a, b in zip(W, c["W"]):
        assert np.abs(a - b).max() < FACTOR_TOL


# ---- the CUDA path ----------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def H():
    return importlib.import_module("pairwise-perturbation_b200.host_api")


@pytest.fixture(scope="module")
def world(H):
    w = H.World(0, solver=0, use_graph=True, workspace_bytes=512 << 20)
    yield w
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [0, 1])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_reference_fixture(H, world, name, solver):
    """solver 0 = Cholesky-type R x R inverse (the default), 1 = pseudo-inverse with SVD semantics (what the reference's
    SVD_solve computes, common.cxx:710-725)."""
    c = load(name)
    lens, R, N, vnorm = c["lens"], c["R"], len(c["lens"]), c["vnorm"]
    world.set(solver=solver, use_graph=True)
    if c["op"] == "hosvd":
        if solver == 1:
            pytest.skip("no R x R solve on the Tucker path")
        ranks = [int(x) for x in c["g"]["ranks"]]
        V = o.make_tensor_r2(lens)
        Vd = H.Tensor.from_numpy(world, V)
        Wd = [H.Matrix(world, lens[i], ranks[i]) for i in range(N)]
        cored = H.Tensor(world, tuple(ranks))
        H.hosvd(world, Vd, cored, Wd, ranks)
        for i in range(N):
            assert proj_err(Wd[i].numpy(), c["W"][i]) < 1e-8
        assert abs(np.linalg.norm(cored.numpy()) - float(c["g"]["core_norm"])) < 1e-10 * vnorm
        for x in [Vd, cored] + Wd:
            x.free()
        return
    if c["op"].startswith("alsTucker"):
        if solver == 1:
            pytest.skip("no R x R solve on the Tucker path")
        V = o.make_tensor_r2(lens)
        Vd = H.Tensor.from_numpy(world, V)
        W0 = [c["g"]["W_init%d" % i] for i in range(N)]
        Wd = [H.Tensor.from_numpy(world, w, matrix=True) for w in W0]
        cored = H.Tensor.from_numpy(world, o.TTMc(V, W0, -1))
        with H.Trace() as t:
            if c["op"] == "alsTucker_DT":
                H.alsTucker_DT(world, Vd, cored, Wd, 1e-10 * vnorm, c["maxiter"], resprint=c["resprint"])
            else:
                H.alsTucker_PP(world, Vd, cored, Wd, 1e-10 * vnorm, c["tol_init"], c["maxiter"],
                               resprint=c["resprint"])
        assert t.events == c["events"]
        check_rows(t.rows, c, tucker=True)
        for i in range(N):
            assert proj_err(Wd[i].numpy(), c["W"][i]) < 1e-7
        for x in [Vd, cored] + Wd:
            x.free()
        world.set(solver=0, use_graph=True)
        return
    V, _ = o.make_tensor_r(lens, R)
    W, G = o.init_factors(lens, R), o.init_grad(lens, R)
    Vd = H.Tensor.from_numpy(world, V)
    Wd = [H.Tensor.from_numpy(world, w, matrix=True) for w in W]
    Gd = [H.Tensor.from_numpy(world, g, matrix=True) for g in G]
    Fd = [H.Matrix(world, w.shape[0], w.shape[1]) for w in W]
    with H.Trace() as t:
        if c["op"] == "alsCP_DT":
            H.alsCP_DT(world, Vd, Wd, Gd, Fd, 1e-10 * vnorm, c["maxiter"], lam=c["lambda_"], resprint=c["resprint"])
        elif c["op"] == "alsCP_PP":
            H.alsCP_PP(world, Vd, Wd, Gd, Fd, 1e-10 * vnorm, c["tol_init"], c["maxiter"], lam=c["lambda_"],
                       ratio_step=c["ratio_step"], resprint=c["resprint"])
        else:
            H.alsCP_PP_partupdate(world, Vd, Wd, Gd, Fd, 1e-10 * vnorm, c["tol_init"], c["maxiter"], lam=c["lambda_"],
                                  ratio_step=c["ratio_step"], update_percentage=c["update_pct"],
                                  resprint=c["resprint"])
    world.set(solver=0, use_graph=True)
    assert t.events == c["events"]
    check_rows(t.rows, c)
    for i in range(N):
        assert np.abs(Wd[i].numpy() - c["W"][i]).max() < FACTOR_TOL
    for x in [Vd] + Wd + Gd + Fd:
        x.free()
