"""Generates the committed golden fixtures from the reference tree's own test data.
Run once in the build container (needs /root/reference):   python tests/golden/make_golden.py

Fixtures (all derived from files under /root/reference/tests/data, the reference's own
known-answer tests for this path):
  lens_ps_400.npy        ps_lensinput[4,4,401] built exactly as tests/test_pixell.py:208-215 does
                         (powspec.read_camb_scalar on test_scalCls.dat), truncated to lmax=400
  unlensed_071123.npz    every 2nd row (incl. both pole rows) of MM_unlensed_071123.fits =
                         alm2map(rand_alm(ps,lmax=400,seed=1)[1:], spin=[0,2]) on the CC 181x360 grid,
                         plus the WCS numbers from the FITS header
  lensed_071123.npz      every 2nd row of MM_lensed_071123.fits = lensing.rand_map(seed=1, lmax=400, geodesic=True) on the
                         same grid (tests/test_pixell.py:351-356): alm2map(deriv) -> offset_by_grad -> alm2map_pos
  offset_071123.npz      every 10th row of MM_offset_{obs_pos,grad,raw_pos}_071123.fits (tests/test_pixell.py:200-206, 333-349):
                         raw_pos = lensing.offset_by_grad(obs_pos, grad, pol=True, geodesic=True)
  pixels_041121.npz      the 36 reference pixels + mean square of MM_041121.pkl['fullsky_10arc_car']
                         for both spectra (rand_map(seed=10,lmax=1500), tests/test_pixell.py:568-580)
"""
import os, sys, pickle, importlib.util
import numpy as np
REF = os.environ.get("PIXELL_REF", "/root/reference")
here = os.path.dirname(os.path.abspath(__file__))

def load_ref_module(name):
	"""Import pixell.<name> from the reference tree without running pixell/__init__ (astropy is absent)."""
	import types
	if "pixell" not in sys.modules:
		pkg = types.ModuleType("pixell"); pkg.__path__ = [os.path.join(REF, "pixell")]
		sys.modules["pixell"] = pkg
	spec = importlib.util.spec_from_file_location("pixell."+name, os.path.join(REF, "pixell", name+".py"))
	mod = importlib.util.module_from_spec(spec); sys.modules["pixell."+name] = mod
	spec.loader.exec_module(mod)
	return mod

def read_fits_primary(fname):
	with open(fname, "rb") as f: raw = f.read()
	hdr = {}
	pos = 0
	while True:
		card = raw[pos:pos+80].decode("ascii"); pos += 80
		key = card[:8].strip()
		if key == "END": break
		if card[8] == "=":
			val = card[10:].split("/")[0].strip()
			hdr[key] = val
	pos = (pos+2879)//2880*2880
	shape = tuple(int(hdr["NAXIS%d" % i]) for i in range(int(hdr["NAXIS"]), 0, -1))
	assert int(hdr["BITPIX"]) == -64
	data = np.frombuffer(raw, ">f8", int(np.prod(shape)), pos).reshape(shape).astype(np.float64)
	return data, hdr

def main():
	utils = load_ref_module("utils"); powspec = load_ref_module("powspec")
	D = os.path.join(REF, "tests", "data")
	ps_cmb, ps_lens = powspec.read_camb_scalar(os.path.join(D, "test_scalCls.dat"))
	ps = np.zeros((4,4,ps_cmb.shape[-1])); ps[0,0] = ps_lens; ps[1:,1:] = ps_cmb
	np.save(os.path.join(here, "lens_ps_400.npy"), ps[:,:,:401])
	m, hdr = read_fits_primary(os.path.join(D, "MM_unlensed_071123.fits"))
	rows = np.arange(0, m.shape[1], 2)
	np.savez(os.path.join(here, "unlensed_071123.npz"), rows=rows, map=m[:,rows],
		shape=np.array(m.shape), crpix=[float(hdr["CRPIX1"]), float(hdr["CRPIX2"])],
		cdelt=[float(hdr["CDELT1"]), float(hdr["CDELT2"])], crval=[float(hdr["CRVAL1"]), float(hdr["CRVAL2"])])
	ml, _ = read_fits_primary(os.path.join(D, "MM_lensed_071123.fits"))
	np.savez(os.path.join(here, "lensed_071123.npz"), rows=rows, map=ml[:, rows])
	r10 = np.arange(0, m.shape[1], 10)
	off = {k: read_fits_primary(os.path.join(D, "MM_offset_%s_071123.fits" % k))[0][:, r10] for k in ("obs_pos", "grad", "raw_pos")}
	np.savez(os.path.join(here, "offset_071123.npz"), rows=r10, **off)
	pk = pickle.load(open(os.path.join(D, "MM_041121.pkl"), "rb"))["fullsky_10arc_car"]
	np.savez(os.path.join(here, "pixels_041121.npz"),
		white_10=pk["white_10"]["refpixels"], white_10_ms=pk["white_10"]["meansquare"],
		constant_dl_1=pk["constant_dl_1"]["refpixels"], constant_dl_1_ms=pk["constant_dl_1"]["meansquare"])
	print("wrote fixtures to", here)

if __name__ == "__main__": main()
