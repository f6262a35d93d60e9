"""oracle/rotate_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU checker; never imported by the product).

rotate_alm (reference pixell/curvedsky.py:717-742 -> ducc0.sht.rotate_alm / healpy.rotate_alm) restated two independent ways:

  rotate_alm_wigner      a'_lm = sum_m' D^l_{m m'}(phi, theta, psi) a_lm',  D^l_{mm'}(a, b, c) = e^{-i m a} d^l_{mm'}(b) e^{-i m' c},
                         Wigner's small d from its explicit factorial sum (Varshalovich 4.3.1 (2)), lmax <= ~40
  rotate_alm_bruteforce  no group theory at all: the field is evaluated with scipy's Y_lm at R^-1 x for the nodes x of an exact
                         Gauss-Legendre x equispaced quadrature (3x3 rotation matrices), and projected back on Y_lm

Convention (pinned by constants the reference publishes): the field is rotated ACTIVELY by R = Rz(phi) Ry(theta) Rz(psi),
f'(x) = f(R^-1 x).  pixell/curvedsky.py:714-716 gives gal -> equ as (psi, theta, phi) = (57.068, 62.871, -167.141) degrees =
(180 deg - l_NCP, 90 deg - dec_NGP, ra_NGP - 360 deg): R must carry the galactic pole z to (ra_NGP, dec_NGP) and the
direction of the celestial pole in galactic coordinates (l_NCP, b = dec_NGP) to z -- checked in tests/test_oracle_basic.py.
pixell/curvedsky.py:578 (prof2alm) uses (0, pi/2 - dec, ra) to move a polar profile to [ra, dec]: the same convention.
alm layout: m-major triangular, real fields (a_{l,-m} = (-1)^m conj a_lm), healpy normalisation.
"""
import math
import numpy as np

def rotmat(psi, theta, phi):
	def rz(a): c, s = np.cos(a), np.sin(a); return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
	def ry(a): c, s = np.cos(a), np.sin(a); return np.array([[c, 0, s], [0, 1.0, 0], [-s, 0, c]])
	return rz(phi) @ ry(theta) @ rz(psi)

def _idx(lmax, l, m): return m*(2*lmax+1-m)//2 + l

def wigner_d(l, beta):
	"""d^l_{m m'}(beta), m, m' = -l..l (array index m + l), by the explicit sum"""
	c, s = math.cos(beta/2), math.sin(beta/2)
	f = [math.factorial(i) for i in range(2*l+2)]
	d = np.zeros((2*l+1, 2*l+1))
	for m in range(-l, l+1):
		for mp in range(-l, l+1):
			pre = math.sqrt(f[l+m]*f[l-m]*f[l+mp]*f[l-mp])
			tot = 0.0
			for k in range(max(0, mp-m), min(l+mp, l-m)+1):
				# Varshalovich / Sakurai: d^l_{m mp} = sum_k (-1)^(k - mp + m) ... cos^(2l - 2k + mp - m) sin^(2k - mp + m)
				tot += (-1)**(k-mp+m)/(f[l+mp-k]*f[k]*f[l-k-m]*f[k-mp+m])*c**(2*l-2*k+mp-m)*s**(2*k-mp+m)
			d[m+l, mp+l] = pre*tot
	return d

def rotate_alm_wigner(alm, lmax, psi, theta, phi):
	alm = np.asarray(alm, dtype=np.complex128)
	out = np.zeros_like(alm)
	for l in range(lmax+1):
		d = wigner_d(l, theta)
		m = np.arange(-l, l+1)
		full = np.zeros(2*l+1, complex)
		for mm in range(0, l+1):
			a = alm[_idx(lmax, l, mm)]
			full[mm+l] = a
			if mm > 0: full[-mm+l] = (-1)**mm*np.conj(a)
		D = np.exp(-1j*m[:, None]*phi)*d*np.exp(-1j*m[None, :]*psi)
		res = D @ full
		for mm in range(0, l+1): out[_idx(lmax, l, mm)] = res[mm+l]
	return out

def _ylm_all(lmax, theta, phi):
	"""Y_lm(theta, phi) for m >= 0 in the m-major triangular order, [nalm, npoint]"""
	try: from scipy.special import sph_harm_y as sy; f = lambda l, m: sy(l, m, theta, phi)
	except ImportError:
		from scipy.special import sph_harm as sh; f = lambda l, m: sh(m, l, phi, theta)
	out = np.zeros(((lmax+1)*(lmax+2)//2, len(theta)), complex)
	for m in range(lmax+1):
		for l in range(m, lmax+1): out[_idx(lmax, l, m)] = f(l, m)
	return out

def synth_points(alm, lmax, theta, phi):
	Y = _ylm_all(lmax, theta, phi)
	w = np.where(np.arange(len(alm)) <= lmax, 1.0, 2.0)      # m = 0 once, m > 0 with the conjugate partner
	return np.real(np.sum((w*alm)[:, None]*Y, 0))

def rotate_alm_bruteforce(alm, lmax, psi, theta, phi):
	x, wq = np.polynomial.legendre.leggauss(lmax+2)
	nphi = 2*lmax+4
	th = np.repeat(np.arccos(x), nphi); ph = np.tile(np.arange(nphi)*2*np.pi/nphi, len(x)); w = np.repeat(wq, nphi)*2*np.pi/nphi
	v = np.stack([np.sin(th)*np.cos(ph), np.sin(th)*np.sin(ph), np.cos(th)])
	u = rotmat(psi, theta, phi).T @ v                          # R^-1 x
	f = synth_points(np.asarray(alm, complex), lmax, np.arccos(np.clip(u[2], -1, 1)), np.arctan2(u[1], u[0]))
	return _ylm_all(lmax, th, ph).conj() @ (w*f)
