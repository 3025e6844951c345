"""INTEGRATION.md option B, compiled: the reference's OWN main-cli.c with the bodies of smvp_csr_compute
(/root/reference/main-cli.c:325) and smvp_tjds_compute (:734) replaced by calls into libsmvp_cuda.

CPU part (build container, /root/reference present): integration/apply_dropin.py splices the bodies into a scratch copy
of the reference's file and gcc builds it against include/smvp_cuda.h -- the drop-in really compiles and binds every
entry point the brief names.  GPU part: the binary built by __graft_entry__.build() (integration/_build/, shipped to the
GPU box like oracle/_ref) reproduces the reference's golden CSR reports through the reference's own main(), loader and
report writer (call sites main-cli.c:1457, :1469)."""
import glob
import importlib.util
import os
import subprocess

import numpy as np
import pytest

import util

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
EXE = os.path.join(REPO, "integration", "_build", "smvp-toolkit-cli-dropin")


def _apply():
    spec = importlib.util.spec_from_file_location("apply_dropin", os.path.join(REPO, "integration", "apply_dropin.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main-cli.c")), reason="the reference is only mounted in the build container")
def test_dropin_splices_and_links(tmp_path):
    mod = _apply()
    out = mod.splice(os.path.join(REF, "main-cli.c"), str(tmp_path / "main-cli.c"))
    text = open(out).read()
    ref = open(os.path.join(REF, "main-cli.c")).read()
    # only the two bodies and one include changed: everything before / between / after them is the reference's text
    assert '#include "smvp_cuda.h"' in text and "smvp_csr_build(" in text and "smvp_tjds_mult(" in text
    assert "qsort(mmImportData" not in text.split("double *smvp_csr_compute(")[1].split("void smvp_cisr_coegen(")[0]
    for keep in ("void smvp_cisr_coegen(", "void generateReportText(", "int main(int argc", "double calcStDevDouble("):
        assert keep in text
    assert len(text) < len(ref) - 12000, "the CPU implementations are gone"
    exe = mod.build(REF, str(tmp_path))
    und = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("smvp_csr_build", "smvp_csr_mult", "smvp_tjds_build", "smvp_tjds_mult", "smvp_time_stats", "smvp_csr_free",
                "smvp_tjds_free"):
        assert " U %s" % sym in und, sym
    assert not os.path.exists(str(tmp_path / "main-cli.dropin.c")), "the spliced copy of the reference's file is not kept"


def _close(y, ref, coo):
    scale = np.abs(coo["val"]).max() if len(coo) else 1.0
    return np.all(np.abs(y - ref) <= 1e-5 * np.abs(ref) + 1e-12 * scale * 600)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ibm32", "curtis54", "pdp08-pg4", "memplus", "pwt"])
def test_dropin_reproduces_golden_reports(tmp_path, name):
    if not os.path.exists(EXE):
        pytest.skip("integration/_build/smvp-toolkit-cli-dropin not built (needs /root/reference at build time)")
    m, n, coo = util.load_sample(name)
    rd = str(tmp_path) + "/"
    # the reference's own option parser and report writer: -c / -t (its --all-algs runs nothing, U1), -d is mandatory (U2)
    r = subprocess.run([EXE, "-c", "-n", "7", "-d", rd, util.sample_path(name)], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Calculating 7 iterations of SMVP CSR." in r.stdout
    rep = util.parse_report(glob.glob(str(tmp_path / "smvp-toolbox_report_CSR_*.txt"))[0])
    assert rep["nnz"] == len(coo) and rep["iters"] == 7 and len(rep["y"]) == m
    assert rep["min"] <= rep["avg"] <= rep["max"]
    if (name, "CSR") in util.GOLDEN_REPORTS:
        gold = util.parse_report(os.path.join(util.GOLDEN, "reports", util.GOLDEN_REPORTS[(name, "CSR")]))
        if name == "pwt":  # the golden pwt report carries the reference's uninitialised row_ptr[0] in y[0] (U3)
            assert _close(rep["y"][1:], gold["y"][1:], coo)
        else:
            assert _close(rep["y"], gold["y"], coo)
    # TJDS through the same binary: full product == CSR; with SMVP_DROPIN_REF_COMPAT=1 the shipped truncated product
    r = subprocess.run([EXE, "-t", "-n", "3", "-d", rd, util.sample_path(name)], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rt = util.parse_report(sorted(glob.glob(str(tmp_path / "smvp-toolbox_report_TJDS_*.txt")))[-1])
    assert _close(rt["y"], rep["y"], coo)
    if (name, "TJDS") in util.GOLDEN_REPORTS:
        for f in glob.glob(str(tmp_path / "smvp-toolbox_report_TJDS_*.txt")):
            os.remove(f)
        env = dict(os.environ, SMVP_DROPIN_REF_COMPAT="1")
        r = subprocess.run([EXE, "-t", "-n", "3", "-d", rd, util.sample_path(name)], capture_output=True, text=True,
                           cwd=str(tmp_path), env=env)
        assert r.returncode == 0
        rt = util.parse_report(glob.glob(str(tmp_path / "smvp-toolbox_report_TJDS_*.txt"))[0])
        gold = util.parse_report(os.path.join(util.GOLDEN, "reports", util.GOLDEN_REPORTS[(name, "TJDS")]))
        assert _close(rt["y"], gold["y"], coo)
