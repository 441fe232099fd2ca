"""Make the reference importable on the GPU box for ``bench.py --impl reference``.

    python tools/install_reference.py            # in the build container (needs /root/reference)

1. tries the contract's offline install (``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse
   --target baseline/_ref /root/reference``) and records why it fails: the reference has no ``setup.py`` /
   ``pyproject.toml`` (it is a directory of scripts), so pip has nothing to build;
2. copies the reference's Python sources (28 ``.py`` files + requirements.txt, ~150 KB; no images, no checkpoints) into the
   git-ignored ``baseline/_ref/`` -- never into history.  ``.gpurunignore`` does not list it, so it travels with the
   ``gpurun`` snapshot.  ``bench.py --impl reference`` imports it with the five stub modules of
   ``tests/golden/make_golden.py`` and falls back to the oracle port (``cpu_baseline.kind = "port"``) when it is absent.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(REF):
        print("no /root/reference here: nothing to do")
        return 1
    os.makedirs(DST, exist_ok=True)
    tmp = "/tmp/fd_ref_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp, ignore=shutil.ignore_patterns("*.pth", "*.jpg", "*.png", ".git"))
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                        "/opt/wheelhouse", "--no-deps", "--target", DST, tmp], capture_output=True, text=True)
    note = (r.stdout + r.stderr).strip().splitlines()[-1:] if r.returncode else ["pip install succeeded"]
    print("pip install:", "exit", r.returncode, "|", " ".join(note))
    n = 0
    for dp, dn, files in os.walk(REF):
        dn[:] = [d for d in dn if d not in (".git", "imgs", "saved_models", "__pycache__")]
        for f in files:
            if f.endswith(".py") or f == "requirements.txt":
                rel = os.path.relpath(os.path.join(dp, f), REF)
                os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                shutil.copyfile(os.path.join(dp, f), os.path.join(DST, rel))
                n += 1
    with open(os.path.join(DST, "INSTALL_NOTE.txt"), "w") as fp:
        fp.write(f"pip exit {r.returncode}: {' '.join(note)}\ncopied {n} source files from {REF}\n")
    print(f"copied {n} files into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
