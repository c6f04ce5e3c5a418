#!/usr/bin/env python
"""Build recipe for ``oracle/_ref/``: the UNMODIFIED reference, byte-compiled.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            # needs /root/reference (the build container)

The reference's hot path is one pure-Python module (``/root/reference/manytor.py``; numpy + stdlib, no
build system), so "compiling it from the sources where they lie" is ``py_compile``: the output is
``oracle/_ref/manytor_ref.bin`` -- CPython bytecode (the content of a .pyc) of the reference file, nothing edited, no source copied
into the repository.  ``oracle/_ref/`` is git-ignored (it stays out of history) but not gpurun-ignored,
so it travels to the GPU box like the built ``.so``; there ``bench.py`` imports it (sourceless import)
to time the reference's own ``Multienv`` loop (test_multi.py:11-34) on the host cores, and
``tests/test_oracle_golden.py`` replays a golden trace through it when it is present.  Without
``/root/reference`` the existing ``oracle/_ref`` is kept; without either, callers fall back to the numpy
port and say so (``cpu_baseline.kind = "port"``).
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(os.environ.get("MANYTOR_REFERENCE", "/root/reference"), "manytor.py")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "manytor_ref.bin")    # CPython bytecode; not named *.pyc, which snapshot tools tend to skip
META = os.path.join(OUT_DIR, "BUILD.json")


def build(verbose: bool = True) -> str | None:
    """-> path of the byte-compiled reference, or None when neither the source nor an earlier build exists."""
    if not os.path.exists(REF_SRC):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    py_compile.compile(REF_SRC, cfile=OUT, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(REF_SRC, "rb") as f:
        digest = hashlib.sha256(f.read()).hexdigest()
    with open(META, "w") as f:
        json.dump({"source": REF_SRC, "sha256": digest, "python": sys.version.split()[0],
                   "recipe": "py_compile.compile(source, cfile='oracle/_ref/manytor_ref.bin')"}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: byte-compiled {REF_SRC} (sha256 {digest[:12]}) -> {OUT}")
    return OUT


def load():
    """Import the byte-compiled reference as module ``manytor`` (None when it was never built)."""
    if not os.path.exists(OUT):
        return None
    import importlib.machinery
    import importlib.util
    loader = importlib.machinery.SourcelessFileLoader("manytor", OUT)
    spec = importlib.util.spec_from_loader("manytor", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def multienv_loop(steps: int, env_shape=(8, 8), obj_number: int = 10, seed: int = 0):
    """The reference's own loop (test_multi.py:11-34 without tqdm/render): Multienv(env_shape, obj_number),
    reset(), `steps` x step(action_sample()).  Returns (env-steps, seconds) or None without oracle/_ref."""
    import time
    import numpy as np
    ref = load()
    if ref is None:
        return None
    np.random.seed(seed)
    me = ref.Multienv(env_shape=env_shape, obj_number=obj_number)
    me.reset()
    t0 = time.perf_counter()
    for _ in range(steps):
        me.step(me.action_sample())
    return me.env_number * steps, time.perf_counter() - t0


if __name__ == "__main__":
    p = build()
    if p is None:
        sys.exit("no reference source and no earlier build")
    r = multienv_loop(5)
    print(f"reference Multienv((8,8),10): {r[0] / r[1]:.1f} env-steps/s on 1 core")
