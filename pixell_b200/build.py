"""Builds pixell_b200/libb200sht.so from csrc/*.cu with nvcc for sm_100a (in-tree, so the
shared object travels with the repository snapshot to the GPU box)."""
import os, subprocess, sys, concurrent.futures

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200sht.so")
SOURCES = ["api.cu", "legendre.cu", "ringfft.cu", "resample.cu", "almops.cu", "fft2d.cu", "tfft.cu", "general.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
	"-Xcompiler", "-fPIC", "-Xptxas", "-v"]

def _newest_input():
	t = 0
	for root in (CSRC, os.path.join(HERE, "..", "include")):
		for f in os.listdir(root):
			t = max(t, os.path.getmtime(os.path.join(root, f)))
	return t

def build(force=False, verbose=False):
	if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_input():
		return OUT
	objdir = os.path.join(HERE, "build"); os.makedirs(objdir, exist_ok=True)
	def compile_one(src):
		obj = os.path.join(objdir, src.replace(".cu", ".o"))
		cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
		r = subprocess.run(cmd, capture_output=True, text=True)
		if r.returncode != 0:
			raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
		return obj, r.stderr
	with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
		res = list(ex.map(compile_one, SOURCES))
	if verbose:
		for obj, log in res: sys.stderr.write(log)
	with open(os.path.join(objdir, "ptxas.log"), "w") as f:
		for obj, log in res: f.write("### %s\n%s\n" % (os.path.basename(obj), log))
	cmd = [NVCC, "-shared", "-o", OUT] + [o for o, _ in res] + ["-gencode", "arch=compute_100a,code=sm_100a"]
	r = subprocess.run(cmd, capture_output=True, text=True)
	if r.returncode != 0: raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
	return OUT

if __name__ == "__main__":
	print(build(force="--force" in sys.argv, verbose=True))
