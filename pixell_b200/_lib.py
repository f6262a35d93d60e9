"""ctypes binding of libb200sht.so (include/b200sht.h).  The product path has no CPU fallback:
if the CUDA library is missing or no device is usable, every call fails loudly."""
import ctypes, os, threading
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_HERE, "libb200sht.so")
_lib = None
_lock = threading.Lock()
_tls = threading.local()

MEM_HOST, MEM_DEVICE = 0, 1
F64, F32 = 0, 1
MODE_STANDARD, MODE_DERIV1 = 0, 1
FFT_C2C, FFT_R2C, FFT_C2R = 0, 1, 2

c_int, c_i64, c_dbl, c_vp, c_cp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_char_p
_i64p, _dblp, _intp = ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl), ctypes.POINTER(c_int)

_PROTOS = {
	"b2_init": ([c_int], c_int),
	"b2_last_error": ([], c_cp),
	"b2_version": ([], c_int),
	"b2_device_synchronize": ([], c_int),
	"b2_launch_count": ([], c_i64),
	"b2_dfma_peak_gflops": ([_dblp], c_int),
	"b2_sht_plan_rings": ([ctypes.POINTER(c_vp), c_int, _dblp, c_i64, c_dbl, c_int, c_i64, _i64p, _dblp, c_int, c_int, _i64p, c_i64], c_int),
	"b2_sht_plan_rings_general": ([ctypes.POINTER(c_vp), c_int, _dblp, _i64p, _dblp, _i64p, _dblp, c_int, c_int, _i64p, c_i64], c_int),
	"b2_sht_plan_2d": ([ctypes.POINTER(c_vp), c_cp, c_int, c_i64, c_dbl, c_int, c_int, c_int, c_int, _i64p, c_i64], c_int),
	"b2_sht_plan_destroy": ([c_vp], None),
	"b2_sht_plan_bytes": ([c_vp], c_i64),
	"b2_synthesis": ([c_vp, c_int, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_vp], c_int),
	"b2_adjoint_synthesis": ([c_vp, c_int, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_vp], c_int),
	"b2_analysis_2d": ([c_vp, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_vp], c_int),
	"b2_adjoint_analysis_2d": ([c_vp, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_vp], c_int),
	"b2_sht_execute_groups": ([c_vp, c_int, c_int, _intp, c_int, c_int, ctypes.POINTER(c_vp), _i64p, ctypes.POINTER(c_vp), _i64p, c_int, c_vp], c_int),
	"b2_sht_last_timing": ([c_vp, _dblp], c_int),
	"b2_alm2leg": ([c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_vp], c_int),
	"b2_leg2alm": ([c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_vp], c_int),
	"b2_theta_weighting": ([c_vp, c_int, c_int, c_int, c_vp, c_vp], c_int),
	"b2_set_leg_variant": ([c_int, c_int], c_int),
	"b2_gridweights": ([c_cp, c_int, _dblp], c_int),
	"b2_alm2cl": ([c_int, c_int, _i64p, c_int, c_vp, c_vp, c_int, c_vp, c_int, c_vp], c_int),
	"b2_lmul": ([c_int, c_int, _i64p, c_int, c_vp, c_int, c_vp, c_int, c_vp], c_int),
	"b2_lmatmul": ([c_int, c_int, c_int, c_int, _i64p, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_int, c_vp], c_int),
	"b2_transpose_alm": ([c_int, c_int, _i64p, c_int, c_vp, c_vp, c_int, c_vp], c_int),
	"b2_transfer_alm": ([c_int, c_int, _i64p, c_i64, c_vp, c_int, c_int, _i64p, c_i64, c_vp, c_int, c_int, c_vp], c_int),
	"b2_rand_alm": ([c_int, c_int, _i64p, c_int, ctypes.c_uint64, c_vp, c_int, c_vp, c_i64, c_int, c_vp], c_int),
	"b2_fft_plan_create": ([ctypes.POINTER(c_vp), c_int, _i64p, _i64p, _i64p, c_int, _intp, c_int, c_int], c_int),
	"b2_fft_execute": ([c_vp, c_vp, c_vp, c_int, c_dbl, c_int, c_vp], c_int),
	"b2_fft_plan_destroy": ([c_vp], None),
	"b2_general_extend": ([c_vp, c_vp, c_int, c_int, c_int, c_i64, c_int, c_vp], c_int),
	"b2_general_scatter": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp], c_int),
	"b2_general_interp": ([c_vp, c_int, c_int, c_vp, c_i64, c_int, c_dbl, c_vp, c_i64, c_vp], c_int),
	"b2_general_spread": ([c_vp, c_int, c_int, c_vp, c_i64, c_int, c_dbl, c_vp, c_i64, c_vp], c_int),
	"b2_general_gather": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp], c_int),
	"b2_general_fold": ([c_vp, c_vp, c_int, c_int, c_int, c_i64, c_int, c_vp], c_int),
	"b2_fourier_filter": ([c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp], c_int),
	"b2_queb_rotate": ([c_vp, c_i64, c_i64, c_i64, c_int, c_int, _dblp, _dblp, c_int, c_int, c_int, c_int, c_vp], c_int),
}

class B200Error(RuntimeError): pass

def lib():
	"""The loaded shared library (loads it on first use)."""
	global _lib
	if _lib is None:
		with _lock:
			if _lib is None:
				if not os.path.exists(LIBPATH):
					raise ImportError("pixell_b200: %s is missing; build it with `python -m pixell_b200.build` "
						"(nvcc, sm_100a).  There is no CPU fallback." % LIBPATH)
				L = ctypes.CDLL(LIBPATH)
				for name, (args, res) in _PROTOS.items():
					f = getattr(L, name)       # AttributeError here means header and library disagree
					f.argtypes = args; f.restype = res
				_lib = L
	return _lib

def last_error():
	return lib().b2_last_error().decode("utf-8", "replace")

def check(rc, exc=B200Error):
	if rc != 0: raise exc(last_error())

def init(device=None):
	"""Bind this thread to a CUDA device (default: PIXELL_B200_DEVICE, LOCAL_RANK, or torch's current device)."""
	if device is None:
		env = os.environ.get("PIXELL_B200_DEVICE", os.environ.get("LOCAL_RANK"))
		if env is not None: device = int(env)
		else:
			device = 0
			try:
				import torch
				if torch.cuda.is_available(): device = torch.cuda.current_device()
			except ImportError: pass
	if getattr(_tls, "device", None) != device:
		check(lib().b2_init(int(device)))
		_tls.device = device
	return device

def is_torch(a):
	return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")

def buffer_info(a):
	"""(pointer, mem kind, numpy dtype) of a numpy array or torch tensor."""
	if is_torch(a):
		import torch
		dt = {torch.float64: np.float64, torch.float32: np.float32, torch.complex128: np.complex128,
			torch.complex64: np.complex64, torch.int64: np.int64}[a.dtype]
		return a.data_ptr(), (MEM_DEVICE if a.is_cuda else MEM_HOST), np.dtype(dt)
	return a.ctypes.data, MEM_HOST, a.dtype

def strides_elems(a):
	if is_torch(a): return tuple(a.stride())
	return tuple(s//a.itemsize for s in a.strides)

def current_stream(a=None):
	"""torch's current CUDA stream for device tensors (so our kernels order with torch work); else the default stream."""
	if a is not None and is_torch(a) and a.is_cuda:
		import torch
		return torch.cuda.current_stream(a.device).cuda_stream
	return None

def as_i64(a):
	return np.ascontiguousarray(np.asarray(a).astype(np.int64))

def p_i64(a): return a.ctypes.data_as(_i64p)
def p_dbl(a): return a.ctypes.data_as(_dblp)
