"""CPU-side checks of the product: the C-ABI library loads and exports every symbol that
include/fora_b200.h declares, its host-side helpers (loader, CSR, parameter derivation) are
bit-exact against the oracle / golden vectors, and it refuses to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import fora_b200 as fb
from helpers import GOLDEN_DIR, ROOT, Graph, Oracle, write_dataset


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fora_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fora_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = fb.lib()
    missing = [name for name in sorted(declared) if not hasattr(L, name)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", fb.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fora_[a-z0-9_]+)", out))
    assert declared <= exported


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fb.ForaError, match="no CUDA device"):
        fb.Engine()


def test_product_does_not_reference_oracle():
    # the oracle is test infrastructure: nothing under fora_b200/ may import, include or link it
    for root, _, files in os.walk(os.path.join(ROOT, "fora_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "fora_oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(root, f)
    out = subprocess.run(["ldd", fb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_host_loader_and_csr_bit_exact():
    gold = np.load(os.path.join(GOLDEN_DIR, "ref_small.npz"))
    n = int(gold["n"])
    d = tempfile.mkdtemp()
    write_dataset(d, n, int(gold["m_decl"]), gold["src"], gold["dst"])
    assert fb.read_attribute(os.path.join(d, "attribute.txt")) == (n, int(gold["m_decl"]))
    src, dst = fb.read_edges(os.path.join(d, "graph.txt"), n)
    keep = gold["src"] != gold["dst"]
    assert np.array_equal(src, gold["src"][keep]) and np.array_equal(dst, gold["dst"][keep])
    op, oc, ip_, ic = fb.csr_from_edges(n, gold["src"], gold["dst"])
    assert np.array_equal(op, gold["out_ptr"]) and np.array_equal(oc, gold["out_col"])
    assert np.array_equal(ip_, gold["in_ptr"]) and np.array_equal(ic, gold["in_col"])
    # id >= n is an error (the reference asserts, graph.h:155-156)
    with open(os.path.join(d, "graph.txt"), "a") as f:
        f.write("%d 1\n" % n)
    with pytest.raises(fb.ForaError):
        fb.read_edges(os.path.join(d, "graph.txt"), n)
    with pytest.raises(fb.ForaError):
        fb.read_attribute(os.path.join(d, "missing.txt"))


def test_host_loader_edge_cases():
    d = tempfile.mkdtemp()
    # empty edge file, ragged whitespace, trailing newline missing
    open(os.path.join(d, "e.txt"), "w").write("")
    s, t = fb.read_edges(os.path.join(d, "e.txt"), 5)
    assert len(s) == 0
    open(os.path.join(d, "r.txt"), "w").write("0 1\n\n  2\t3 \r\n4 4\n1   0")
    s, t = fb.read_edges(os.path.join(d, "r.txt"), 5)
    assert s.tolist() == [0, 2, 1] and t.tolist() == [1, 3, 0]
    op, oc, ip_, ic = fb.csr_from_edges(5, s, t)
    assert op.tolist() == [0, 1, 2, 3, 3, 3] and oc.tolist() == [1, 0, 3]
    assert ip_.tolist() == [0, 1, 2, 2, 3, 3] and ic.tolist() == [1, 0, 2]


def test_host_settings_bit_exact():
    gold = np.load(os.path.join(GOLDEN_DIR, "ref_small.npz"))
    n, m, eps = int(gold["n"]), int(gold["m_decl"]), float(gold["eps"])
    for opt in (0, 1):
        for w in ("fora", "fora_topk", "montecarlo", "bippr", "fwdpush"):
            ref = gold["setting_%s_opt%d" % (w, opt)]
            rmax, omega = fb.setting(w, n, m, eps, opt=opt)
            if w != "montecarlo":
                assert rmax == ref[0]
            if w != "fwdpush":
                assert omega == ref[1]
    # named shapes of SURVEY.md section 8 (values quoted there to 4 digits)
    rmax, omega = fb.setting("fora", 4847571, 68993773, 0.5)
    assert abs(rmax / 3.936e-9 - 1) < 1e-3 and abs(omega / 7.798e8 - 1) < 1e-3
    rmax, omega = fb.setting("fora", 4847571, 68993773, 0.5, opt=1)
    assert abs(rmax / 4.919e-9 - 1) < 1e-3


def test_synthetic_generator_is_deterministic_and_shaped():
    a = fb.synth_edges(20000, 200000, seed=42)
    b = fb.synth_edges(20000, 200000, seed=42)
    c = fb.synth_edges(20000, 200000, seed=43)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and not np.array_equal(a[0], c[0])
    assert (a[0] != a[1]).all() and a[0].min() >= 0 and a[1].max() < 20000
    op, oc, ip_, ic = fb.csr_from_edges(20000, *a)
    deg = np.diff(op)
    assert len(oc) == 200000 and (deg == 0).sum() >= 0.03 * 20000 and deg.max() > 20 * deg.mean()
    # the C ABI CSR equals the numpy restatement used to feed the oracle
    g = Graph(20000, a[0], a[1])
    assert np.array_equal(op, g.out_ptr) and np.array_equal(oc, g.out_col) and np.array_equal(ip_, g.in_ptr) and np.array_equal(ic, g.in_col)


def test_parallel_loader_matches_sequential_semantics():
    # many chunks (the loader cuts the file per 64 KB per thread), pairs straddling chunk boundaries, ragged whitespace
    rng = np.random.default_rng(5)
    n, ne = 50000, 400000
    src = rng.integers(0, n, ne)
    dst = rng.integers(0, n, ne)
    dst[::97] = src[::97]  # self loops
    seps = np.array([" ", "\t", "\n", "  ", " \n", "\r\n"])
    toks = np.empty(2 * ne, dtype=object)
    toks[0::2] = src.astype(str)
    toks[1::2] = dst.astype(str)
    sp = seps[rng.integers(0, len(seps), 2 * ne)]
    text = "".join(t + s for t, s in zip(toks, sp))
    d = tempfile.mkdtemp()
    path = os.path.join(d, "graph.txt")
    open(path, "w").write(text)
    s2, d2 = fb.read_edges(path, n)
    keep = src != dst
    assert np.array_equal(s2, src[keep]) and np.array_equal(d2, dst[keep])
    open(path, "w").write(text + " 7")  # dangling last token: fscanf stops, the pair is never formed
    s3, d3 = fb.read_edges(path, n)
    assert np.array_equal(s3, s2) and np.array_equal(d3, d2)
