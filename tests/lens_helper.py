"""Test infrastructure: the host-side geometry between the two engine calls of pixell's curved-sky lensing
(reference pixell/lensing.py:552-621 offset_by_grad / offset_by_grad_helper, geodesic=True), restated with numpy so
that the reference's lensing golden (tests/test_pixell.py:351-356) can be replayed through the CUDA path.  Pinned by
tests/test_oracle_golden.py::test_offset_by_grad_golden against the reference's own offset goldens."""
import numpy as np

def offset_by_grad(ipos, grad):
	"""ipos[{dec,ra},...], grad[{d/ddec, d/dra / cos dec},...] -> opos[{dec, ra, psi},...]: move every point along the
	great circle in the direction of grad by |grad| and return the rotation psi of the local polarisation basis."""
	dec, ra = ipos[0].reshape(-1), ipos[1].reshape(-1)
	g = np.array(grad, dtype=np.float64).reshape(grad.shape[0], -1)   # every row enters the step length, as in the reference
	g[0] = -g[0]                                                     # zenith convention: theta grows southwards
	g[:, np.all(g == 0, 0)] = 1e-20
	d = np.sqrt(np.sum(g**2, 0)); g = g/d
	th = np.pi/2-dec
	cd, sd, ct, st = np.cos(d), np.sin(d), np.cos(th), np.sin(th)
	oct_ = cd*ct - sd*st*g[0]
	ost = np.sqrt(1-oct_**2)
	ophi = ra + np.arcsin(sd*g[1]/ost)
	with np.errstate(divide="ignore", invalid="ignore"):
		A = g[1]/(sd*ct/st + g[0]*cd)
	n = g[0] + g[1]*A
	den = 1+A**2
	cg, sg = 2*n**2/den-1, 2*n*(g[1]-g[0]*A)/den
	out = np.array([np.pi/2-np.arccos(oct_), ophi, np.arctan2(sg, cg)])
	return out.reshape((3,)+ipos.shape[1:])
