"""TEST INFRASTRUCTURE: runs the reference's own code (oracle/_ref/, built by `make -C oracle ref` from the unmodified
sources under /root/reference against oracle/ctf_standin/ctf.hpp) on explicit inputs and parses what it prints.

Only tests/, tests/golden/make_golden_ref.py and bench.py's CPU legs may import this.  Nothing here reads
/root/reference at run time: the binaries in oracle/_ref/ are self-contained and travel to the GPU box with the repo.
"""
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ref_driver"))


def _f(a):
    return np.asarray(a, dtype=np.float64).ravel(order="F")


_ROW = re.compile(r"\[(?:iter|sweeps)\]=\s+(\S+)\s+\[(?:gradnorm|diffnorm)\]\s+(\S+)\s+\[tol\]\s+\S+\s+\[pp_update\]\s+(\d)"
                  r"\s+\[(?:diffV|residual)\]\s+(\S+)")
_EVT = re.compile(r"(DT|pairwise perturbation) starts from (\d+)")


def parse_stdout(text):
    """rows (iter, gradnorm, pp_update, diffV) as the reference prints them (als_CP.cxx:193-196, 17 digits when the
    stream precision was raised) and the switching markers (als_CP.cxx:1110,1118; als_Tucker.cxx:933,941)."""
    rows = [(float(m.group(1)), float(m.group(2)), int(m.group(3)), float(m.group(4))) for m in _ROW.finditer(text)]
    events = [("DT" if m.group(1) == "DT" else "PP", int(m.group(2))) for m in _EVT.finditer(text)]
    return rows, events


def parse_csv(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            p = line.strip().split(",")
            if len(p) == 7 and not p[0].startswith("["):
                rows.append((float(p[1]), float(p[2]), int(p[4]), float(p[5])))
    return rows


def run_driver(op, V, W=None, grad=None, R=None, ranks=None, timeout=600, **kw):
    """Calls oracle/_ref/ref_driver.  V: ndarray (any memory order; written in first-index-fastest order), W / grad:
    lists of s_i x R arrays.  kw: tol, tol_init, maxiter, lambda_, ratio_step, update_pct, resprint, bench.
    Returns dict(rows, events, stdout, W, grad, core, files)."""
    assert available(), "oracle/_ref/ref_driver is missing: run `make -C oracle ref` where /root/reference exists"
    lens = V.shape
    N = len(lens)
    with tempfile.TemporaryDirectory() as td:
        _f(V).tofile(os.path.join(td, "V.bin"))
        cmd = [os.path.join(REF_DIR, "ref_driver"), "-op", op, "-lens", ",".join(str(x) for x in lens),
               "-V", os.path.join(td, "V.bin"), "-out", os.path.join(td, "o"), "-csv", os.path.join(td, "o.csv")]
        if R is None and W is not None:
            R = W[0].shape[1]
        if R is not None:
            cmd += ["-rank", str(R)]
        if ranks is not None:
            cmd += ["-ranks", ",".join(str(x) for x in ranks)]
        if W is not None:
            np.concatenate([_f(w) for w in W]).tofile(os.path.join(td, "W.bin"))
            cmd += ["-W", os.path.join(td, "W.bin")]
        if grad is not None:
            np.concatenate([_f(g) for g in grad]).tofile(os.path.join(td, "G.bin"))
            cmd += ["-grad", os.path.join(td, "G.bin")]
        names = dict(lambda_="-lambda")
        for k, v in kw.items():
            cmd += [names.get(k, "-" + k), repr(float(v)) if isinstance(v, float) else str(int(v))]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError("ref_driver failed (%d): %s\n%s" % (res.returncode, " ".join(cmd), res.stderr[-2000:]))
        out = dict(stdout=res.stdout, files={})
        out["rows_stdout"], out["events"] = parse_stdout(res.stdout)
        out["rows"] = parse_csv(os.path.join(td, "o.csv"))
        tucker = op == "hosvd" or op.startswith("alsTucker")
        rk = list(ranks) if (tucker and ranks is not None) else [R] * N

        def split(path):
            flat = np.fromfile(path)
            res_, off = [], 0
            for i in range(N):
                n = lens[i] * rk[i]
                res_.append(flat[off:off + n].reshape((lens[i], rk[i]), order="F").copy())
                off += n
            return res_

        for fn in os.listdir(td):
            if not fn.startswith("o.") or not fn.endswith(".bin"):
                continue
            key = fn[2:-4]
            if key in ("W", "grad", "normalized"):
                out[key if key != "normalized" else "normalized"] = split(os.path.join(td, fn))
            else:
                out["files"][key] = np.fromfile(os.path.join(td, fn))
        if "core" in out["files"]:
            out["core"] = out["files"]["core"].reshape(tuple(rk), order="F")
        return out


def run_cli(binary, args, fills=None, cwd=None, timeout=600):
    """Runs one of the reference's own mains (test_ALS, pp_bench, run, test_decomposition) built against the
    stand-in.  `fills`: list of (seed, id) pairs consumed by successive fill_random calls."""
    exe = os.path.join(REF_DIR, binary)
    assert os.path.exists(exe), exe + " is missing"
    env = dict(os.environ)
    if fills:
        env["CTF_STANDIN_FILLS"] = ",".join("%d:%d" % (s, i) for s, i in fills)
    with tempfile.TemporaryDirectory() as td:
        res = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, env=env,
                             cwd=cwd or td, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (binary, res.returncode, res.stderr[-2000:]))
    rows, events = parse_stdout(res.stdout)
    return dict(stdout=res.stdout, rows=rows, events=events)
