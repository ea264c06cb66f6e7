"""Fourier conventions of the reference (myfft.py:15-52):
   Fourier1D  (t->w): f(w_j) = dt  sum_i f(t_i) exp(+2 pi i ij/N)      = ifft(a) * N dt
   iFourier1D (w->t): f(t_i) = dw/2pi sum_j f(w_j) exp(-2 pi i ij/N)   = fft(a) / (N dt)
Arrays living on the GPU (torch tensors [N][cols]) are transformed by the library's column
FFT; small host arrays go through numpy (analysis helper, not on the stepping path)."""
import numpy as N


class myfft:
    def __init__(self, dt, n):
        self.dt = dt
        self.N = n
        self.dw = 2 * N.pi / dt / n

    def _dev(self, a, sign, scale):
        import torch
        from . import _lib
        z = a.to(torch.complex128).contiguous().clone()
        n, cols = z.shape[0], z[0].numel()
        work = torch.empty(_lib.load().pymd_fft_work_bytes(n, cols), dtype=torch.uint8, device=z.device)
        _lib.call("pymd_fft_columns", z.data_ptr(), n, cols, sign, work.data_ptr(), _lib.stream_ptr())
        return z * scale

    def Fourier1D(self, a):
        if len(a) != self.N:
            raise ValueError("MyFFT.Fourier1D: array length error!")
        if hasattr(a, "is_cuda") and a.is_cuda:
            return self._dev(a, +1, self.dt)               # ifft * N dt = (unscaled +) * dt
        return (2. * N.pi / self.dw) * N.fft.ifft(a, axis=0)

    def iFourier1D(self, a):
        if len(a) != self.N:
            raise ValueError("MyFFT.iFourier1D: array length error!")
        if hasattr(a, "is_cuda") and a.is_cuda:
            return self._dev(a, -1, 1.0 / (self.N * self.dt))
        return (self.dw / 2 / N.pi) * N.fft.fft(a, axis=0)
