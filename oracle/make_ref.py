"""Recipe for oracle/_ref: an importable copy of the reference's OWN implementation of the hot path.

The reference is pure Python (no build system, nothing to compile); its "build" for the purposes of the CPU baseline is
a copy of the modules the MCEM path imports (and of the three evaluate scripts that call it), taken from where they lie under /root/reference into oracle/_ref/
(git-ignored: the files never enter this repository's history; not gpurun-ignored, so they travel to the GPU box like
a compiled reference would).  bench.py --impl reference then times the reference's own MCEM_M1 / MCEM_M2 classes
(kind "reference"); without oracle/_ref it falls back to the oracle port (kind "port").

    python oracle/make_ref.py          # run in the build container; __graft_entry__.build() calls it when /root/reference exists

python/processing/stft.py is NOT copied: it imports librosa, which is absent from this image; the reference arm uses the
numpy restatement oracle/stft_oracle.py for the two ends of the path and says so in its JSON line.
"""
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["python/__init__.py", "python/models/__init__.py", "python/models/mcem.py", "python/models/models.py",
         "python/models/distributions.py", "python/processing/__init__.py", "python/processing/target.py", "python/metrics.py",
         # the three evaluate scripts: tests/test_gpu_scripts.py runs their process_utt UNMODIFIED against the drop-in
         # modules of guided-vae-nmf_b200/python (the drop-in boundary)
         "scripts/evaluate_M1.py", "scripts/evaluate_M2_ibm.py", "scripts/evaluate_M2_vad.py"]


def make_ref():
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        else:                                               # a package directory without an __init__.py in the reference
            open(dst, "a").close()
    return True


if __name__ == "__main__":
    print("oracle/_ref written" if make_ref() else "no /root/reference here: oracle/_ref left as it is")
