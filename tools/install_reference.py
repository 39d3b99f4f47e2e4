"""Installs the UNMODIFIED reference under the git-ignored ``baseline/_ref/`` so that it travels to the GPU box with
the gpurun snapshot (like ``oracle/_ref``): the GPU tests and ``bench.py --impl reference`` then run the reference's
own ``models.architectures.KPFCNN`` / ``datasets`` code instead of a restatement. Run in the build container:

    python tools/install_reference.py

The contract's installer (``pip install --target baseline/_ref /root/reference``) is tried first; the reference is a
script tree without setup.py / pyproject.toml, so pip refuses it, and the Python packages the path needs are then
placed by a plain file copy. Nothing under ``baseline/_ref`` is ever committed (``.gitignore``), and no product code
reads it: ``oracle/ref_harness.py`` (test infrastructure) is its only consumer.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WEASAL_REF_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ["models", "datasets", "kernels", "utils"]
SCRIPTS = ["train_Vaihingen3D_PseudoLabel.py", "train_DALES_PseudoLabel.py", "train_Vaihingen3D_WeakLabel.py",
           "train_DALES_WeakLabel.py", "test_models.py", "LICENSE"]


def main():
    if not os.path.isdir(os.path.join(REF, "models")):
        print(f"reference sources not found at {REF}: keeping {DST} as it is")
        return 0
    os.makedirs(DST, exist_ok=True)
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                          "--find-links", "/opt/wheelhouse", "--target", DST, REF], capture_output=True, text=True)
    how = "pip"
    if pip.returncode != 0 or not os.path.isdir(os.path.join(DST, "models")):
        how = "copy (pip: " + (pip.stderr.strip().splitlines() or ["failed"])[-1][:160] + ")"
        for p in PACKAGES:
            shutil.rmtree(os.path.join(DST, p), ignore_errors=True)
            shutil.copytree(os.path.join(REF, p), os.path.join(DST, p),
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for s in SCRIPTS:
            if os.path.exists(os.path.join(REF, s)):
                shutil.copy2(os.path.join(REF, s), os.path.join(DST, s))
    with open(os.path.join(DST, "INSTALLED_FROM.txt"), "w") as f:
        f.write(f"source: {REF}\nmethod: {how}\n")
    print(f"installed the reference into {DST} by {how}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
