"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Packs the UNMODIFIED reference sources of the RSSFormer path (`/root/reference/RSSFormer-TIP2023/**/*.py`) into ONE archive,
`oracle/_ref/rssformer_reference.zip`, so that the reference itself can travel to the GPU box (where `/root/reference` does not
exist) for `bench.py --impl reference`, the `gpu_eager_baseline` leg and the module-surgery parity test.  `oracle/_ref/` is
git-ignored (no reference source enters the history) but not gpurun-ignored.  Python imports straight from the archive
(zipimport): nothing is unpacked, nothing is edited.

    python -m oracle.build_ref          # also run by __graft_entry__.build() when /root/reference is present
"""
import os
import zipfile

SRC = os.environ.get("RSS_REFERENCE_SRC", "/root/reference/RSSFormer-TIP2023")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ZIP = os.path.join(OUT_DIR, "rssformer_reference.zip")


def build(force=False):
    """-> path of the archive, or None when the reference tree is not present (GPU box: the prebuilt file is used)"""
    if not os.path.isdir(os.path.join(SRC, "module", "baseline")):
        return ZIP if os.path.exists(ZIP) else None
    files = []
    for root, _dirs, names in os.walk(SRC):
        for n in sorted(names):
            if n.endswith(".py"):
                files.append(os.path.join(root, n))
    files.sort()
    newest = max(os.path.getmtime(f) for f in files)
    if not force and os.path.exists(ZIP) and os.path.getmtime(ZIP) >= newest:
        return ZIP
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = ZIP + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        seen = set()
        for f in files:
            rel = os.path.relpath(f, SRC)
            d = os.path.dirname(rel)
            parts = d.split(os.sep) if d else []
            for i in range(1, len(parts) + 1):      # explicit directory entries: zipimport needs them for packages without __init__.py
                dn = "/".join(parts[:i]) + "/"
                if dn not in seen:
                    seen.add(dn)
                    z.writestr(zipfile.ZipInfo(dn), b"")
            z.write(f, rel)
    os.replace(tmp, ZIP)
    return ZIP


if __name__ == "__main__":
    print(build(force=True))
