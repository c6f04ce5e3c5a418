#!/usr/bin/env python
"""Build libmanytor_b200.so of another revision next to the working tree's, for tools/ab.py:

    python tools/build_variant.py <git-ref> <name>      ->  build/variants/<name>.so

`build/` is git-ignored but travels to the GPU box with the snapshot.  The revision must have the same
C ABI version as the working tree's bindings (manytor_b200/_lib.py checks it at load)."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ref, name = sys.argv[1], sys.argv[2]
    out_dir = os.path.join(ROOT, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        wt = os.path.join(tmp, "wt")
        subprocess.run(["git", "-C", ROOT, "worktree", "add", "--detach", wt, ref], check=True, capture_output=True)
        try:
            subprocess.run([sys.executable, os.path.join(wt, "manytor_b200", "build.py")], check=True, capture_output=True)
            shutil.copy(os.path.join(wt, "manytor_b200", "lib", "libmanytor_b200.so"), os.path.join(out_dir, name + ".so"))
        finally:
            subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", wt], check=False, capture_output=True)
    print(os.path.join(out_dir, name + ".so"))


if __name__ == "__main__":
    main()
