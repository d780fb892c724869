#!/usr/bin/env python
"""Stage the UNMODIFIED reference hot-path files into ``oracle/_ref/`` (git-ignored, travels with gpurun).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is a Python program: its sampling hot path imports with two stub
modules (``oracle/ref_runner.py``).  This recipe copies, byte for byte, exactly the files that path needs from
``/root/reference`` (read-only, present in the build container only) so that the GPU box can run the real reference
(a) as the same-device parity comparator of ``tests/test_gpu_reference.py``, and (b) as the baseline arm of
``bench.py --impl reference`` / ``gpu_reference``.  Nothing under ``hierdiff_b200/`` reads ``oracle/_ref``.

    python oracle/stage_ref.py            # copies when /root/reference exists; otherwise verifies the staged copy

A manifest with the sha256 of every staged file is written next to them; ``verify()`` checks it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("HD_REFERENCE_ROOT", "/root/reference")

# (path relative to the reference root).  Stage 1: the coarse-grained sampler; stage 2: the fine-grained decoder's
# equivariant layer.
FILES = [
    "endiffusion/train_module/__init__.py",
    "endiffusion/train_module/diffusion_qm9.py",
    "endiffusion/models/__init__.py",
    "endiffusion/models/utils.py",
    "endiffusion/models/noise_model.py",
    "endiffusion/models/distributions.py",
    "endiffusion/models/module/__init__.py",
    "endiffusion/models/module/en_dynamics.py",
    "endiffusion/models/layers/__init__.py",
    "endiffusion/models/layers/egnn_new.py",
    "endiffusion/equivariant_diffusion/__init__.py",
    "endiffusion/equivariant_diffusion/utils.py",
    "endiffusion/loss/__init__.py",
    "endiffusion/loss/criterion.py",
    "endiffusion/dataset/__init__.py",
    "endiffusion/dataset/datasets_statistics.py",
    "endiffusion/conf/model/ddpmgblur.yaml",
    "endiffusion/conf/analyze/GEOM.yaml",
    "endiffusion/conf/sample.yaml",
    "endiffusion/conf/sample/default.yaml",
    "models/__init__.py",
    "models/egnn/gcl.py",
    "models/egnn/egnn_new.py",
    "models/egnn/utils.py",
    "models/edge_denoise.py",
    "models/flows/__init__.py",
    "models/flows/utils.py",
    "data_utils/data_diffuse.py",     # only its breadth-first helpers are ever executed (make_golden_stage2.import_edge_denoise)
    "conf/model/edge_denoise.yaml",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage():
    """Copy FILES from the reference tree; returns the manifest."""
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DEST, rel)
        if not os.path.exists(src):
            if os.path.basename(rel) == "__init__.py":     # namespace-style directory in the reference: nothing to copy
                continue
            raise FileNotFoundError(src)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify():
    """True when oracle/_ref holds every staged file with the recorded hash."""
    path = os.path.join(DEST, "MANIFEST.json")
    if not os.path.exists(path):
        return False
    with open(path) as f:
        manifest = json.load(f)["files"]
    return all(os.path.exists(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == h
               for rel, h in manifest.items())


def available():
    return verify()


def ensure():
    """Stage when the reference tree is present (build container); otherwise the staged copy must verify."""
    if os.path.isdir(SRC):
        stage()
    return verify()


if __name__ == "__main__":
    ok = ensure()
    print("oracle/_ref:", "ok" if ok else "MISSING (no /root/reference here and nothing staged)")
    sys.exit(0 if ok else 1)
