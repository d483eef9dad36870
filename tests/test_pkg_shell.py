"""The R package shell (pkg/): written without R in the image, so these host checks pin what can be pinned -- the nvcc
rules cover every kernel source, and each .Call site passes exactly the number of arguments the glue's entry point
takes (the reference's arities, src/spNNGP.h:9-45, src/nn.h:18,25,38)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_makevars_lists_every_kernel_source():
    mk = open(os.path.join(ROOT, "pkg", "src", "Makevars")).read()
    assert "-gencode arch=compute_100a,code=sm_100a" in mk and "$(NVCC)" in mk
    listed = set(re.findall(r"(\w+\.cu)\b", mk.split("CU_OBJS")[0]))
    have = {os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "spnngp_b200", "csrc", "*.cu"))}
    assert listed == have, (sorted(have - listed), sorted(listed - have))


def _call_arity(text, symbol):
    """number of arguments after the symbol name in each .Call("symbol", ...) of `text`"""
    out = []
    for m in re.finditer(r'\.Call\("%s"' % symbol, text):
        i = text.index("(", m.start()); depth = 0; args = 0; j = i
        while True:
            ch = text[j]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            elif ch == "," and depth == 1:
                args += 1
            j += 1
        out.append(args)
    return out


def test_call_sites_match_the_glue_arities():
    r_src = "".join(open(p).read() for p in glob.glob(os.path.join(ROOT, "pkg", "R", "*.R")))
    r_src = "\n".join(ln for ln in r_src.split("\n") if not ln.lstrip().startswith("##"))  # code only
    glue = open(os.path.join(ROOT, "spnngp_b200", "rglue", "spnngp_glue.cpp")).read()
    for sym in ("rNNGP", "cNNGP", "rNNGPPredict", "mkNNIndx", "mkNNIndxCB"):
        m = re.search(r"\bSEXP %s\(([^)]*)\)" % sym, glue)
        assert m, sym
        want = m.group(1).count("SEXP")
        got = _call_arity(r_src, sym)
        assert got and all(g == want for g in got), (sym, want, got)


def test_namespace_exports_the_entry_points():
    ns = open(os.path.join(ROOT, "pkg", "NAMESPACE")).read()
    assert "useDynLib(spNNGP, .registration = TRUE)" in ns
    for line in ("export(spNNGP)", "export(spConjNNGP)", 'S3method("predict", "NNGP")'):
        assert line in ns
    r_src = "".join(open(p).read() for p in glob.glob(os.path.join(ROOT, "pkg", "R", "*.R")))
    for fn in ("spNNGP <- function(", "spConjNNGP <- function(", "predict.NNGP <- function("):
        assert fn in r_src
