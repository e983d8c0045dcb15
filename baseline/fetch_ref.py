"""Materialise the unmodified reference under baseline/_ref/ (git-ignored, NOT gpurun-ignored).

    python baseline/fetch_ref.py [/root/reference]

/root/reference exists only in the build container.  The reference is a plain Python tree without
a setup.py (nothing to `pip install`), so the "install" is a byte-for-byte copy of its *.py files
into baseline/_ref/, which travels to the GPU box with the snapshot like the built .so files and
never enters git history.  Nothing under baseline/_ref/ is imported by the product
(online-continual-learning_b200/); it is used by
  * bench.py --impl reference      (the reference's own agents on the host cores),
  * bench.py's reference_gpu figure (the same agents on cuda:0 -- the >=10x denominator),
  * tests/test_gpu_dropin.py        (reference nn.Module + optimizer -> install() -> train_learner).
A manifest with the sha256 of every copied file is written next to the copy so that a test can
prove the copy is unmodified.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SKIP_DIRS = {'.git', '__pycache__', 'config', 'config_CVPR'}


def fetch(src='/root/reference', quiet=False):
    if not os.path.isdir(src):
        if not quiet:
            print('fetch_ref: %s not present (GPU box?) -- keeping existing %s' % (src, DST))
        return os.path.isdir(DST)
    manifest = {}
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, src)
        for f in files:
            if not f.endswith('.py'):
                continue
            s = os.path.join(root, f)
            d = os.path.join(DST, rel, f) if rel != '.' else os.path.join(DST, f)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            with open(s, 'rb') as fh:
                data = fh.read()
            manifest[os.path.normpath(os.path.join(rel, f))] = hashlib.sha256(data).hexdigest()
            if not os.path.exists(d) or open(d, 'rb').read() != data:
                shutil.copyfile(s, d)
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': src, 'files': manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print('fetch_ref: %d files -> %s' % (len(manifest), DST))
    return True


def verify():
    """True when every file listed in the manifest is present and byte-identical to what was copied."""
    path = os.path.join(DST, 'MANIFEST.json')
    if not os.path.exists(path):
        return False
    with open(path) as fh:
        files = json.load(fh)['files']
    for rel, digest in files.items():
        p = os.path.join(DST, rel)
        if not os.path.exists(p) or hashlib.sha256(open(p, 'rb').read()).hexdigest() != digest:
            return False
    return True


if __name__ == '__main__':
    ok = fetch(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    sys.exit(0 if ok and verify() else 1)
