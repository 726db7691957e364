"""Consumer side of the hot path: evaluation of the eyebox bin tensor.

Mirrors ``AR_system_evaluation_functions.evaluation`` of the reference
(/root/reference/AR_system_evaluation_functions.py:45-163) and the runner's efficiency numbers
(gpu_ray_tracing_pro_fullColor.py:186-192).  The heavy part -- the pupil-mask sums over the
864 MB bin tensor (reference lines 68-109; 5.5 s of NumPy at the default size) and the per-cell
totals -- runs on the GPU (``wgrt_eval_pupil_sums``); the remaining arithmetic works on
[3, FoV_y, FoV_x, 7, 8] arrays and stays in NumPy, following the reference line by line.

Parity: pupil sums, U_fov, U_EB and output_image are checked against the reference function
itself (tests/golden/eval.npz, rel. tol. 1e-5: the reference sums float32 in NumPy's pairwise
order, the kernel in its own order).  ``delta_e`` needs ``colour-science`` in the reference
(absent here and un-versioned there): CIE Lab / CIEDE2000 are restated from the published
formulas and are "parity unpinned".
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _capi

__all__ = ["evaluation", "evaluation_device", "eval_params", "finish_metrics", "pupil_sums", "efficiency_per_colour",
           "linearize_srgb", "apply_srgb_gamma", "normalize_brightness_without_changing_color"]

# reference lines 47-57
M = np.array([[1.67430115, -0.76582385, -0.06172232],
              [-0.12551154, 1.47840695, -0.04124377],
              [-0.01826868, -0.13098157, 1.61444037]])
M_XYZ = np.array([[6.424000e-01, 1.891400e-01, 2.511000e-01],
                  [2.650000e-01, 8.849624e-01, 7.390000e-02],
                  [4.999999e-05, 3.693564e-02, 1.528100e+00]])
# D65 white (CIE 1931 2 deg): XYZ of the D65 spectrum scaled to Y = 100, and the Lab reference white
# from the chromaticity (0.3127, 0.3290) that colour-science uses by default
_XYZ_D65_SPD = np.array([95.047, 100.0, 108.883])
_WHITE_XY = (0.3127, 0.3290)


def linearize_srgb(image_srgb):
    return np.where(image_srgb <= 0.04045, image_srgb / 12.92, ((image_srgb + 0.055) / 1.055) ** 2.4)


def apply_srgb_gamma(image_linear):
    return np.where(image_linear <= 0.0031308, image_linear * 12.92, 1.055 * (image_linear ** (1 / 2.4)) - 0.055)


def normalize_brightness_without_changing_color(img_srgb_float):
    """Reference lines 18-43 (HSV value stretch); OpenCV when present, else the same maths in NumPy."""
    img = np.asarray(img_srgb_float, dtype=np.float32)
    try:
        import cv2
        hsv = cv2.cvtColor(img, cv2.COLOR_RGB2HSV)
        h, s, v = cv2.split(hsv)
        max_v = np.max(v)
        if max_v > 0:
            v = v / max_v
        return cv2.cvtColor(cv2.merge([h, s, v]), cv2.COLOR_HSV2RGB)
    except ImportError:
        v = img.max(axis=-1)
        max_v = v.max()
        return img / max_v if max_v > 0 else img


def pupil_sums(matrix_EB: np.ndarray, mask_size: int = 30, step_y: int = 8, step_x: int = 12
               ) -> Tuple[np.ndarray, np.ndarray]:
    """GPU: disc-pupil sums at the sampled eye positions (reference lines 68-109) and per-cell totals.

    Returns ``(matrix_eye_perceive [L, Yf, Xf, n_epy, n_epx], cell_sums [L, Yf, Xf])`` as float32.
    """
    lib = _capi.load_library()
    eb = np.ascontiguousarray(matrix_EB, dtype=np.float32)
    if eb.ndim != 5:
        raise ValueError("matrix_EB must be [L, FoV_y, FoV_x, EBy, EBx]")
    L, Yf, Xf, EBy, EBx = eb.shape
    n_epy = (EBy - mask_size) // step_y + 1 if EBy >= mask_size else 0
    n_epx = (EBx - mask_size) // step_x + 1 if EBx >= mask_size else 0
    out = np.zeros((L, Yf, Xf, n_epy, n_epx), dtype=np.float32)
    cells = np.zeros((L, Yf, Xf), dtype=np.float32)
    _capi.check(lib.wgrt_eval_pupil_sums_host(eb.ctypes.data, L, Yf, Xf, EBy, EBx, mask_size, step_y, step_x,
                                              out.ctypes.data, cells.ctypes.data), lib)
    return out, cells


def efficiency_per_colour(matrix_EB: np.ndarray, num_rays: int, num_iter: int) -> np.ndarray:
    """gpu_ray_tracing_pro_fullColor.py:186-192: sum over the eyebox and FoV of each wavelength's bins,
    / num_rays / num_iter * 3.  Index 0 = blue (465 nm) ... 2 = red (630 nm)."""
    _, cells = pupil_sums(matrix_EB)
    return cells.astype(np.float64).sum(axis=(1, 2)) / num_rays / num_iter * 3


def _xyz_to_lab(xyz, white):
    t = xyz / white
    d = 6.0 / 29.0
    f = np.where(t > d ** 3, np.cbrt(t), t / (3 * d * d) + 4.0 / 29.0)
    return np.stack((116 * f[..., 1] - 16, 500 * (f[..., 0] - f[..., 1]), 200 * (f[..., 1] - f[..., 2])), axis=-1)


def _delta_e_2000(lab1, lab2):
    """CIEDE2000 (Sharma, Wu, Dalal 2005), kL = kC = kH = 1."""
    L1, a1, b1 = lab1[..., 0], lab1[..., 1], lab1[..., 2]
    L2, a2, b2 = lab2[..., 0], lab2[..., 1], lab2[..., 2]
    C1, C2 = np.hypot(a1, b1), np.hypot(a2, b2)
    Cm = 0.5 * (C1 + C2)
    G = 0.5 * (1 - np.sqrt(Cm ** 7 / (Cm ** 7 + 25.0 ** 7)))
    a1p, a2p = (1 + G) * a1, (1 + G) * a2
    C1p, C2p = np.hypot(a1p, b1), np.hypot(a2p, b2)
    h1p = np.degrees(np.arctan2(b1, a1p)) % 360
    h2p = np.degrees(np.arctan2(b2, a2p)) % 360
    dLp, dCp = L2 - L1, C2p - C1p
    dh = h2p - h1p
    dh = np.where(C1p * C2p == 0, 0.0, np.where(dh > 180, dh - 360, np.where(dh < -180, dh + 360, dh)))
    dHp = 2 * np.sqrt(C1p * C2p) * np.sin(np.radians(dh) / 2)
    Lm, Cpm = 0.5 * (L1 + L2), 0.5 * (C1p + C2p)
    hsum = h1p + h2p
    hm = np.where(C1p * C2p == 0, hsum,
                  np.where(np.abs(h1p - h2p) <= 180, hsum / 2, np.where(hsum < 360, (hsum + 360) / 2, (hsum - 360) / 2)))
    T = (1 - 0.17 * np.cos(np.radians(hm - 30)) + 0.24 * np.cos(np.radians(2 * hm)) +
         0.32 * np.cos(np.radians(3 * hm + 6)) - 0.20 * np.cos(np.radians(4 * hm - 63)))
    dth = 30 * np.exp(-(((hm - 275) / 25) ** 2))
    Rc = 2 * np.sqrt(Cpm ** 7 / (Cpm ** 7 + 25.0 ** 7))
    Sl = 1 + 0.015 * (Lm - 50) ** 2 / np.sqrt(20 + (Lm - 50) ** 2)
    Sc, Sh = 1 + 0.045 * Cpm, 1 + 0.015 * Cpm * T
    Rt = -np.sin(np.radians(2 * dth)) * Rc
    return np.sqrt((dLp / Sl) ** 2 + (dCp / Sc) ** 2 + (dHp / Sh) ** 2 + Rt * (dCp / Sc) * (dHp / Sh))


def _lab_white():
    x_w, y_w = _WHITE_XY
    return np.array([x_w / y_w, 1.0, (1 - x_w - y_w) / y_w]) * 100.0


def eval_params(scale: float) -> "_capi.WgrtEvalParams":
    """The colour constants of ``evaluation()`` (reference lines 47-63, 112-116) as the device kernel takes
    them (``wgrt_eval_params_t``); ``scale`` multiplies the raw pupil sums (1 / (num_rays_per_FoV * num_iter),
    RUN:197)."""
    prm = _capi.WgrtEvalParams()
    prm.scale = float(scale)
    white = _lab_white()
    w_rgb = np.linalg.inv(M) @ linearize_srgb(np.ones(3))
    lab_d65 = _xyz_to_lab(_XYZ_D65_SPD / _XYZ_D65_SPD[1] * 100.0, white)
    for dst, src in ((prm.white_rgb, w_rgb), (prm.M, M.ravel()), (prm.M_xyz, M_XYZ.ravel()), (prm.white_xyz, white),
                     (prm.lab_d65, lab_d65)):
        for i, v in enumerate(src):
            dst[i] = float(v)
    return prm


def finish_metrics(metrics: np.ndarray, n_pix: int, n_epy: int, n_epx: int):
    """``(delta_e, U_fov, U_EB)`` from the per-eye-position reductions of ``wgrt_eval_metrics`` (reference
    lines 148-160: the averages over eye positions)."""
    m = np.asarray(metrics, dtype=np.float64).reshape(n_epy * n_epx, _capi.WGRT_EVAL_NUM)
    n_ep = n_epy * n_epx
    delta_e = float(np.sum(m[:, 0] / n_pix) / n_ep)
    has_zero = m[:, 4] > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(has_zero, 0.0, m[:, 1] / m[:, 2])
    U_fov = float(np.sum(ratio) / n_ep)
    u_eb = np.where(has_zero, 0.0, m[:, 3] / n_pix)
    U_EB = 0 if np.max(u_eb) == 0 else float(np.min(u_eb) / np.max(u_eb))
    return delta_e, U_fov, U_EB


def evaluation_device(matrix_eye_perceive_raw: np.ndarray, scale: float, return_image: bool = True):
    """``evaluation()`` lines 110-160 on the GPU (``wgrt_eval_metrics``) from RAW pupil sums
    [3, Yf, Xf, n_epy, n_epx] and the normalisation ``scale``.  Returns ``delta_e, U_fov, U_EB, output_image``
    (``output_image`` None unless requested)."""
    lib = _capi.load_library()
    p = np.ascontiguousarray(matrix_eye_perceive_raw, dtype=np.float32)
    if p.ndim != 5 or p.shape[0] != 3:
        raise ValueError("matrix_eye_perceive must be [3, FoV_y, FoV_x, n_epy, n_epx]")
    _, Yf, Xf, n_epy, n_epx = p.shape
    metrics = np.zeros((n_epy * n_epx, _capi.WGRT_EVAL_NUM), dtype=np.float64)
    image = np.zeros((Yf, Xf, 3, n_epy, n_epx), dtype=np.float32) if return_image else None
    prm = eval_params(scale)
    _capi.check(lib.wgrt_eval_metrics_host(p.ctypes.data, Yf, Xf, n_epy, n_epx, C.byref(prm), metrics.ctypes.data,
                                           image.ctypes.data if return_image else None), lib)
    return (*finish_metrics(metrics, Yf * Xf, n_epy, n_epx), image)


def evaluation(matrix_EB, matrix_eye_perceive: Optional[np.ndarray] = None):
    """Same signature and return tuple as the reference: ``delta_e, U_fov, U_EB, output_image``.

    ``matrix_EB`` is the NORMALISED bin tensor the runner passes (RUN:197-198).  ``matrix_eye_perceive``
    may be supplied when the pupil sums were already reduced on the device.
    """
    M_inv = np.linalg.inv(M)
    x_w, y_w = _WHITE_XY
    white = np.array([x_w / y_w, 1.0, (1 - x_w - y_w) / y_w]) * 100.0
    LAB_D65 = _xyz_to_lab(_XYZ_D65_SPD / _XYZ_D65_SPD[1] * 100.0, white)

    n_lambda, n_FOVy, n_FOVx, n_eby, n_ebx = matrix_EB.shape
    if matrix_eye_perceive is None:
        matrix_eye_perceive, _ = pupil_sums(matrix_EB)             # reference lines 68-109, on the GPU
    n_epy, n_epx = matrix_eye_perceive.shape[3:]

    # pure white input through the inverse sensor matrix (reference lines 112-121)
    img_linear = linearize_srgb(np.zeros((n_FOVy, n_FOVx, 3)) + 1.0)
    wavelength_image = (M_inv @ img_linear.reshape(-1, 3).T).T.reshape(n_FOVy, n_FOVx, 3)[..., None, None]
    adjusted = wavelength_image * np.flip(np.transpose(matrix_eye_perceive, (1, 2, 0, 3, 4)), axis=2)
    output_image = np.empty_like(adjusted)
    delta_e, U_fov = 0.0, 0.0
    U_EB = np.zeros((n_epy, n_epx))
    for i in range(n_epy):
        for j in range(n_epx):
            px = adjusted[:, :, :, i, j].reshape(-1, 3)
            srgb = np.clip((M @ px.T).T.reshape(n_FOVy, n_FOVx, 3), 0, 1)
            output_image[:, :, :, i, j] = normalize_brightness_without_changing_color(apply_srgb_gamma(srgb))
            xyz = (M_XYZ @ px.T).T.reshape(n_FOVy, n_FOVx, 3)
            Y = xyz[:, :, 1]
            xyz_norm = xyz / np.maximum(Y, 1e-10)[..., None] * 100
            lab = _xyz_to_lab(xyz_norm, white)
            lab[Y == 0] = 0
            delta_e += np.mean(_delta_e_2000(lab, LAB_D65))
            if np.any(Y == 0):
                U_EB[i, j] = 0
            else:
                U_fov += np.min(Y) / np.max(Y)
                U_EB[i, j] = np.mean(Y)
    delta_e = delta_e / n_epx / n_epy
    U_fov = U_fov / n_epx / n_epy
    U_EB = 0 if np.max(U_EB) == 0 else np.min(U_EB) / np.max(U_EB)
    return delta_e, U_fov, U_EB, output_image
