"""TEST / BASELINE INFRASTRUCTURE — recipe that places the UNMODIFIED reference's hot-path Python packages under
`oracle/_ref/` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so, and stays out of history).

    python oracle/make_ref.py            # copies from /root/reference (or $FVGN_REFERENCE_ROOT) when it is present

The reference is a directory of Python scripts (no build step): "building" it is copying FVGN-src/main/{FVGN_model,utils,
dataset,Extract_mesh}/*.py byte for byte.  `oracle/ref_loader.py` imports them from /root/reference when that exists (the build
container) and from `oracle/_ref/` otherwise (the GPU box), through the same shims.  Consumers: `bench.py --impl reference`
(the CPU arm, `cpu_baseline.kind = "reference"`), bench.py's `torch_eager_cuda` leg (the reference's own PyTorch path on the
B200) and the `-m gpu` drop-in tests that feed the reference's Data_Pool objects to the CUDA path.  Nothing under
`finite-volume-graph-network_b200/` may import it."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PACKAGES = ("FVGN_model", "utils", "dataset", "Extract_mesh")


def make(src_root=None, verbose=True):
    src_root = src_root or os.environ.get("FVGN_REFERENCE_ROOT", "/root/reference")
    src = os.path.join(src_root, "FVGN-src", "main")
    if not os.path.isfile(os.path.join(src, "FVGN_model", "FVGN.py")):
        if verbose:
            print(f"[make_ref] reference not present at {src}: keeping whatever is in {DST}")
        return os.path.isdir(os.path.join(DST, "FVGN-src", "main", "FVGN_model"))
    out = os.path.join(DST, "FVGN-src", "main")
    manifest = {}
    for pkg in PACKAGES:
        os.makedirs(os.path.join(out, pkg), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(src, pkg))):
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(src, pkg, f), os.path.join(out, pkg, f))
                manifest[f"{pkg}/{f}"] = hashlib.sha256(open(os.path.join(src, pkg, f), "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"[make_ref] {len(manifest)} reference files -> {out}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
