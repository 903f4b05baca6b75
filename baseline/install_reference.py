#!/usr/bin/env python
"""Install the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` (it travels to the GPU box with the gpurun
snapshot; ``/root/reference`` does not exist there).

    python baseline/install_reference.py [--force]

1. The task's recipe is tried first:
       pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
   The reference is a script collection without setup.py / pyproject.toml, so pip answers "Directory is not
   installable" (recorded in baseline/_ref/INSTALL.txt and DESIGN.md).
2. Fallback = what an install of a pure-Python project amounts to: its first-party sources (model/, utils.py, train.py,
   eval.py, predict_video.py, dataloader.py, configs/) are copied byte for byte, plus the *.py files of
   submodules/pycocoevalcap (eval.py imports them at module level).  The rest of submodules/ (340 MB of offline feature
   extractors and Java metric jars, none of it on the hot path) stays behind.

Nothing under baseline/_ref is tracked by git and nothing in the product imports it: it is the reference arm of bench.py
(``--impl reference``, ``gpu_reference``) and the target of tools/run_reference.py on the GPU box.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("VCT_REFERENCE_SRC", "/root/reference")
ITEMS = ["model", "configs", "utils.py", "train.py", "eval.py", "predict_video.py", "dataloader.py", "README.md", "LICENSE"]


def main():
    if not os.path.isdir(SRC):
        if os.path.isdir(os.path.join(DST, "model")):
            print("baseline/_ref already present; source tree not mounted here")
            return 0
        print(f"{SRC} not mounted and baseline/_ref absent: nothing to install", file=sys.stderr)
        return 1
    if os.path.isdir(DST) and "--force" not in sys.argv and os.path.isfile(os.path.join(DST, "INSTALL.txt")):
        print("baseline/_ref up to date (use --force to reinstall)")
        return 0
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                        "/opt/wheelhouse", "--target", DST, SRC], capture_output=True, text=True)
    note = [f"pip install --target baseline/_ref {SRC}: rc={r.returncode}", (r.stderr or r.stdout).strip().splitlines()[-1]
            if (r.stderr or r.stdout).strip() else ""]
    if r.returncode != 0:
        note.append("fallback: first-party python sources copied unmodified")
        digest = hashlib.sha256()
        for it in ITEMS:
            s, d = os.path.join(SRC, it), os.path.join(DST, it)
            if os.path.isdir(s):
                shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            elif os.path.isfile(s):
                shutil.copy2(s, d)
        # the metric package eval.py imports at module level (train.py:8 -> eval.py:11-15): python files only -- the Java
        # jars / models next to them (130 MB) are only needed to SCORE captions, which stays with the reference's scripts
        coco_src = os.path.join(SRC, "submodules", "pycocoevalcap")
        if os.path.isdir(coco_src):
            def only_py(d, names):
                return [n for n in names if not (os.path.isdir(os.path.join(d, n)) or n.endswith(".py"))] + \
                       [n for n in names if n in ("__pycache__", "example")]
            shutil.copytree(coco_src, os.path.join(DST, "submodules", "pycocoevalcap"), ignore=only_py)
            note.append("submodules/pycocoevalcap: *.py only (import-time dependency of eval.py)")
        for root, _dirs, files in sorted(os.walk(DST)):
            for f in sorted(files):
                if f.endswith((".py", ".json")):
                    with open(os.path.join(root, f), "rb") as fh:
                        digest.update(fh.read())
        note.append("sha256(py+json) " + digest.hexdigest())
    with open(os.path.join(DST, "INSTALL.txt"), "w") as f:
        f.write("\n".join(note) + "\n")
    print("\n".join(note))
    return 0


if __name__ == "__main__":
    sys.exit(main())
