#!/usr/bin/env python
"""Regenerates tests/golden/ from the read-only reference checkout.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_fixtures.py [/root/reference]

What is taken, and why (SURVEY.md section 8c lists these as the only fixtures that pin the
hot path at the boundary of the un-vendored likely/cosmo dependencies):

* demo/rmu.{theory,data,icov}, demo/multi.theory, demo/multi_{0,1,2}.{data,cov}
    per-bin model known-answers and a chi-square known-answer (demo/rmu_to_bao.ini,
    demo/multi_to_bao.ini).
* models/<name>.{0,2,4}.dat and <name>_matterpower.dat for the five template families the
    BASELINE configs and the demo inis use (inputs, copied verbatim: the loaders are tested
    on the real on-disk format).
* data/BOSSDR11QSOLyaF.{data,cov,scan}: the only published chi-square surface that can be
    reproduced without the missing cosmology dependency (config/BOSSDR11QSOLyaF.ini).
    The .cov text (2.5 MB) is stored gzip-compressed.
* data/BOSSDR9LyaF.{data,cov,scan} (config/BOSSDR9LyaF.ini: quasar observing coordinates) and
    data/BOSSDR9LyaFXi.{data,cov,scan} (config/BOSSDR9LyaFXi.ini: multipole-binned data,
    data-format = comoving-multipole): the other two published surfaces whose inputs ship.
* the .ini recipes of the hot-path configs, with their relative paths rewritten to the
    fixture tree, comments dropped.

Nothing here is source code of the reference; these are its data files.
"""
import gzip
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

MODEL_FAMILIES = ["DR9LyaMocks", "DR9LyaMocksSB", "DR9LyaMocksNLeffPk",
                  "DR9LyaMocksLCDM", "DR9LyaMocksLCDMSB"]
MATTERPOWER = ["DR9LyaMocksLCDM", "DR9LyaMocksLCDMSB"]
DEMO = ["rmu.theory", "rmu.data", "rmu.icov", "multi.theory",
        "multi_0.data", "multi_0.cov", "multi_1.data", "multi_1.cov",
        "multi_2.data", "multi_2.cov"]
DATA = ["BOSSDR11QSOLyaF.data", "BOSSDR11QSOLyaF.scan", "BOSSDR9LyaF.data",
        "BOSSDR9LyaF.scan", "BOSSDR11LyaF.data", "BOSSDR9LyaFXi.data", "BOSSDR9LyaFXi.scan"]
DATA_GZ = ["BOSSDR11QSOLyaF.cov", "BOSSDR9LyaF.cov", "BOSSDR9LyaFXi.cov"]
INIS = [("config/BOSSDR9LyaF.ini", "BOSSDR9LyaF.ini"),
        ("config/BOSSDR9LyaF_k.ini", "BOSSDR9LyaF_k.ini"),
        ("config/BOSSDR11LyaF_k.ini", "BOSSDR11LyaF_k.ini"),
        ("config/BOSSDR11QSOLyaF_k.ini", "BOSSDR11QSOLyaF_k.ini"),
        ("config/BOSSDR11QSOLyaF.ini", "BOSSDR11QSOLyaF.ini"),
        ("config/BOSSDR11LyaF_fft.ini", "BOSSDR11LyaF_fft.ini"),
        ("config/BOSSDR9LyaFXi.ini", "BOSSDR9LyaFXi.ini"),
        ("demo/rmu_to_bao.ini", "rmu_to_bao.ini"),
        ("demo/multi_to_bao.ini", "multi_to_bao.ini")]


def copy(src, dst):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    os.chmod(dst, 0o644)


def strip_ini(text):
    """Drops comments/blank lines; rewrites ../models ../data ../demo to the fixture tree."""
    out = []
    for line in text.splitlines():
        s = line.strip()
        if not s or s.startswith("#"):
            continue
        s = re.sub(r"\.\./(models|data|demo)/", r"../\1/", s)
        out.append(s)
    return "\n".join(out) + "\n"


def main():
    for fam in MODEL_FAMILIES:
        for ell in (0, 2, 4):
            name = "%s.%d.dat" % (fam, ell)
            copy(os.path.join(REF, "models", name), os.path.join(OUT, "models", name))
    for fam in MATTERPOWER:
        name = fam + "_matterpower.dat"
        copy(os.path.join(REF, "models", name), os.path.join(OUT, "models", name))
    for name in DEMO:
        copy(os.path.join(REF, "demo", name), os.path.join(OUT, "demo", name))
    for name in DATA:
        copy(os.path.join(REF, "data", name), os.path.join(OUT, "data", name))
    for name in DATA_GZ:
        with open(os.path.join(REF, "data", name), "rb") as f:
            raw = f.read()
        dst = os.path.join(OUT, "data", name + ".gz")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with gzip.GzipFile(dst, "wb", mtime=0) as g:
            g.write(raw)
    for src, dst in INIS:
        with open(os.path.join(REF, src)) as f:
            text = strip_ini(f.read())
        p = os.path.join(OUT, "config", dst)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(text)
    print("fixtures written under", OUT)


if __name__ == "__main__":
    main()
