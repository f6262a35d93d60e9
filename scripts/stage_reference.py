#!/usr/bin/env python
"""Stage the UNMODIFIED pure-Python files of the reference (pixell/*.py and tests/test_pixell.py) under baseline/_ref/
(git-ignored; it travels to the GPU box with the repository snapshot), so that tests/test_reference_*.py can import
the reference's own curvedsky.py / enmap.py / fft.py and run the reference's own test functions on top of this
repository's engine (SURVEY.md 8b last row).  Nothing is edited: the files are byte-for-byte copies, checked by hash.
  python scripts/stage_reference.py [/root/reference]"""
import hashlib, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")

def stage(ref="/root/reference"):
	if not os.path.isdir(os.path.join(ref, "pixell")): return None
	os.makedirs(os.path.join(DEST, "pixell"), exist_ok=True); os.makedirs(os.path.join(DEST, "tests"), exist_ok=True)
	manifest = []
	for sub, names in (("pixell", sorted(f for f in os.listdir(os.path.join(ref, "pixell")) if f.endswith(".py"))), ("tests", ["test_pixell.py"])):
		for f in names:
			src = os.path.join(ref, sub, f); dst = os.path.join(DEST, sub, f)
			shutil.copyfile(src, dst)
			manifest.append("%s  %s/%s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), sub, f))
	with open(os.path.join(DEST, "MANIFEST.sha256"), "w") as fh: fh.write("\n".join(manifest)+"\n")
	return DEST

if __name__ == "__main__":
	d = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
	print("staged into %s" % d if d else "reference tree not present: nothing staged")
