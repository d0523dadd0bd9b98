#!/usr/bin/env python
"""Installs the UNMODIFIED reference (iyioon/NYPC-Yacht-Auction, pure Python) into baseline/_ref so that it travels to
the GPU box with the repo snapshot (baseline/_ref is git-ignored, not gpurun-ignored) and can be timed there on the host
cores next to the GPU numbers (bench.py: cpu_baseline.reference_python, --impl reference).

The reference ships no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` refuses it
("not installable").  As the build contract allows, the install is made from a scratch copy under /tmp that only ADDS a
packaging stub; no reference file is edited and nothing is copied into the tracked tree.

    python baseline/install_ref.py [/root/reference]
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")

STUB = '''from setuptools import setup, find_namespace_packages
import glob, os
setup(name="nypc-yacht-auction-reference", version="0",
      py_modules=[os.path.splitext(p)[0] for p in glob.glob("*.py") if p != "setup.py"],
      packages=find_namespace_packages(include=["yacht", "yacht.*"]))
'''


def install(src="/root/reference"):
    if not os.path.isdir(src):
        return "reference sources not present at %s" % src
    if os.path.exists(os.path.join(TARGET, "yacht", "YachtGame.py")):
        return "already installed"
    tmp = tempfile.mkdtemp(prefix="ya_ref_")
    try:
        copy = os.path.join(tmp, "src")
        shutil.copytree(src, copy)
        for d, _, files in os.walk(copy):
            os.chmod(d, 0o755)
            for f in files:
                os.chmod(os.path.join(d, f), 0o644)
        with open(os.path.join(copy, "setup.py"), "w") as f:
            f.write(STUB)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, copy]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            return "pip failed: " + (res.stdout + res.stderr)[-400:]
        return "installed"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(install(*sys.argv[1:2]))
