"""GPU test of the drop-in command line: smvp-toolkit-cli --all-algs on the reference's sample matrices,
report files compared with the reference's golden reports (BASELINE.json configs[0]/[1])."""
import glob
import json
import os
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "smvp-toolkit_b200", "lib", "smvp-toolkit-cli")


def run_cli(args, cwd):
    r = subprocess.run([CLI] + args, capture_output=True, text=True, cwd=cwd)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def close_to_report(y, ref, coo):
    scale = np.abs(coo["val"]).max() if len(coo) else 1.0
    return np.all(np.abs(y - ref) <= 1e-5 * np.abs(ref) + 1e-12 * scale * 600)


@pytest.mark.parametrize("name", util.SAMPLES)
def test_all_algs_reports(tmp_path, name):
    m, n, coo = util.load_sample(name)
    out = run_cli(["--all-algs", "-n", "25", "--json", "-d", str(tmp_path), util.sample_path(name)], str(tmp_path))
    assert "[START]\tExecuting smvp-toolbox-cli v0.6.4" in out and "[STOP]\tExit smvp-toolbox v0.6.4" in out
    assert "Non-zero numbers contained in matrix: \x1b[0m%d" % len(coo) in out
    assert "Calculating 25 iterations of SMVP CSR." in out and "Calculating 25 iterations of SMVP TJDS." in out
    js = [json.loads(ln) for ln in out.splitlines() if ln.startswith("{")]
    assert [j["alg"] for j in js] == ["CSR", "TJDS"] and all(j["iters"] == 25 and j["avg_ms"] > 0 for j in js)
    csr = glob.glob(str(tmp_path / "smvp-toolbox_report_CSR_*.txt"))
    tjds = glob.glob(str(tmp_path / "smvp-toolbox_report_TJDS_*.txt"))
    assert len(csr) == 1 and len(tjds) == 1
    rc, rt = util.parse_report(csr[0]), util.parse_report(tjds[0])
    for rep, alg in ((rc, "CSR"), (rt, "TJDS")):
        assert rep["header"] == "Execution results for smvp-toolbox v.0.6.4, %s algorithm" % alg
        assert rep["nnz"] == len(coo) and rep["iters"] == 25 and len(rep["y"]) == m
        assert rep["min"] <= rep["avg"] <= rep["max"] and abs(rep["total"] - 25 * rep["avg"]) <= 1e-4 * rep["total"]
    # the full TJDS product equals the CSR product (the reference's shipped TJDS does not: U4)
    assert close_to_report(rt["y"], rc["y"], coo)
    if (name, "CSR") in util.GOLDEN_REPORTS:
        gold = util.parse_report(os.path.join(util.GOLDEN, "reports", util.GOLDEN_REPORTS[(name, "CSR")]))
        assert close_to_report(rc["y"], gold["y"], coo)


@pytest.mark.parametrize("name", ["ibm32", "curtis54", "memplus", "pwt"])
def test_ref_compat_reproduces_golden_tjds_reports(tmp_path, name):
    m, n, coo = util.load_sample(name)
    run_cli(["-t", "--ref-compat", "--tjds-variant=deterministic", "-n", "3", "-d", str(tmp_path), util.sample_path(name)],
            str(tmp_path))
    rep = util.parse_report(glob.glob(str(tmp_path / "smvp-toolbox_report_TJDS_*.txt"))[0])
    gold = util.parse_report(os.path.join(util.GOLDEN, "reports", util.GOLDEN_REPORTS[(name, "TJDS")]))
    assert close_to_report(rep["y"], gold["y"], coo)
    if name != "memplus":  # pattern matrices: small integers, the %g text is exact
        assert rep["y_text"] == gold["y_text"]


def test_default_report_dir_is_cwd(tmp_path):
    run_cli(["-c", "-n", "2", util.sample_path("pdp08-pg4")], str(tmp_path))  # no -d: the reference crashes here (U2)
    rep = util.parse_report(glob.glob(str(tmp_path / "smvp-toolbox_report_CSR_*.txt"))[0])
    assert rep["y"].tolist() == [6, 21, 1, 7, 14, 7]


def test_cisr_generator_option(tmp_path):
    """`-g -s 4`: the .coe image on stdout equals the unmodified reference's (tests/golden/cisr)."""
    import gzip

    want = gzip.open(os.path.join(util.GOLDEN, "cisr", "curtis54_s4.coe.gz"), "rt").read()
    out = run_cli(["-g", "-s", "4", util.sample_path("curtis54")], str(tmp_path))
    a = out.index("\n;*********************************************")
    b = out.index("03ffffffff;") + len("03ffffffff;\n\n")
    assert out[a:b] == want
    assert "Converting loaded content to CISR format." in out
