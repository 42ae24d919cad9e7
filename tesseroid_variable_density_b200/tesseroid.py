r"""
Forward model the gravitational fields of tesseroids (spherical prisms) whose density varies
with depth -- B200 (sm_100a) implementation behind the call surface of
``tesseroid_density.tesseroid`` (pinga-lab/tesseroid-variable-density,
code/tesseroid_density/tesseroid.py).

Drop-in surface (same names, argument meaning, defaults, units and error behaviour):

* :func:`potential`, :func:`gx`, :func:`gy`, :func:`gz` (tesseroid.py:327,385,443,501);
* the gradient tensor :func:`gxx` ... :func:`gzz` listed in the reference's docstring
  (tesseroid.py:56-65) but implemented only upstream (Fatiando 0.5); Eotvos, default ratio 8;
* module constants ``RATIO_V, RATIO_G, STACK_SIZE, DELTA`` (tesseroid.py:100-103);
* :func:`_density_based_discretization` (tesseroid.py:237), called directly by two reference
  scripts (number-of-tesseroids.py:77,127; density-based-discretization.py:49).

Coordinate system as in the reference: x -> North, y -> East, z -> up, except ``gz`` which is
computed with z -> Down.  Units: SI for the potential, mGal for gx/gy/gz, Eotvos for the tensor.

Differences that the hardware imposes:

* densities are numbers or density objects (:mod:`.density`), not arbitrary callables;
* ``njobs`` is the number of GPUs the computation points are sharded over
  (the reference's ``multiprocessing`` point chunking, tesseroid.py:179-195, maps 1:1);
  results do not depend on it.  ``pool`` is accepted for signature compatibility and ignored
  apart from the reference's consistency assertion.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib
from .constants import SI2MGAL, SI2EOTVOS, MEAN_EARTH_RADIUS, G
from .density import as_kind_params, NPAR
from .model import FlatModel, flatten_model

RATIO_V = 1
RATIO_G = 2.5
RATIO_GG = 8
STACK_SIZE = 200
DELTA = 1e-2

_WARNING_MSG = (
    "Stopped dividing a tesseroid because it's horizontal dimensions would be " +
    "below the minimum numerical threshold (1e-6 degrees). " +
    "Will compute without division. Cannot guarantee the accuracy of " +
    "the solution.")


def _check_input(lon, lat, height, model, ratio, njobs, pool):
    """Same assertions as the reference's ``_check_input`` (tesseroid.py:106-122)."""
    assert lon.shape == lat.shape == height.shape, \
        "Input coordinate arrays must have same shape"
    assert ratio > 0, "Invalid ratio {}. Must be > 0.".format(ratio)
    assert njobs > 0, "Invalid number of jobs {}. Must be > 0.".format(njobs)
    if njobs == 1:
        assert pool is None, "njobs should be number of processes in the pool"
    return np.zeros_like(lon)


def _convert_coords(lon, lat, height):
    """
    Degrees -> radians, sine/cosine of latitude, height -> radius, with the SAME NumPy calls
    as the reference (tesseroid.py:125-139) so that the values handed to the kernels are
    bit-identical to the ones the reference hands to its Cython kernels.
    """
    lon = np.radians(lon)
    lat = np.radians(lat)
    sinlat = np.sin(lat)
    coslat = np.cos(lat)
    radius = MEAN_EARTH_RADIUS + height
    return lon, sinlat, coslat, radius


def _warm_devices_async(njobs, devices):
    """Start creating the CUDA contexts of the devices a multi-GPU call will use while the model
    is still being flattened (ctypes releases the GIL); returns the thread to join, or None."""
    ids = list(devices) if devices is not None else list(range(int(njobs)))
    if len(ids) < 2:
        return None
    import threading
    lib = _lib.load()
    arr = (C.c_int * len(ids))(*[int(d) for d in ids])
    t = threading.Thread(target=lib.tvd_warm_devices, args=(len(ids), arr))
    t.start()
    return t


def _check_prepared_args(model, dens, delta):
    """A PreparedModel has its densities and radial discretisation baked in: `dens` and a
    different `delta` cannot be honoured, and silently ignoring them would give the result of
    another model."""
    if not isinstance(model, PreparedModel):
        return
    if dens is not None:
        raise ValueError("dens= cannot override the densities of a PreparedModel; build it with "
                         "PreparedModel(model, dens=...)")
    if delta is not DELTA and delta != model.delta and not (delta is None and model.delta is None):
        raise ValueError("delta={!r} differs from the PreparedModel's delta={!r}; build it with "
                         "PreparedModel(model, delta=...)".format(delta, model.delta))


class PreparedModel(object):
    """
    A model flattened and preprocessed on the GPU (density-based radial discretisation done,
    piece table resident).  Field functions accept it in place of ``model``, which skips the
    flattening and the K1 kernel when several fields are computed on the same model.
    """

    def __init__(self, model, dens=None, delta=DELTA, device=0):
        flat = flatten_model(model, dens)
        self.flat = flat
        self.delta = delta
        lib = _lib.load()
        handle = C.c_void_p()
        d = float("nan") if delta is None else float(delta)
        rc = lib.tvd_model_create(flat.size, _lib.ptr_f64(flat.bounds), _lib.ptr_i32(flat.kind),
                                  _lib.ptr_f64(flat.params), d, int(device), C.byref(handle))
        _lib.check(rc, "tvd_model_create")
        self._handle = handle
        self._lib = lib

    @property
    def handle(self):
        return self._handle

    @property
    def n_tesseroids(self):
        return self.flat.size

    @property
    def n_pieces(self):
        return int(self._lib.tvd_model_num_pieces(self._handle))

    def pieces(self):
        """``(top, bottom, parent)`` arrays of the flattened radial split list."""
        n = self.n_pieces
        top, bottom = np.empty(n), np.empty(n)
        parent = np.empty(n, dtype=np.int32)
        rc = self._lib.tvd_model_get_pieces(self._handle, _lib.ptr_f64(top), _lib.ptr_f64(bottom),
                                            _lib.ptr_i32(parent))
        _lib.check(rc, "tvd_model_get_pieces")
        return top, bottom, parent

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.tvd_model_free(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def forward_raw(field, lon_rad, sinlat, coslat, radius, prepared, ratio, njobs=1, devices=None,
                return_stats=False, leaf_counts=False):
    """
    One call of the C ABI: the field in kernel units (field / G), exactly what the reference's
    Cython functions accumulate (``_tesseroid.<field>``, _tesseroid.pyx:204-213).

    Returns ``result`` or ``(result, stats[, counts])``; ``counts`` has shape
    ``(n_tesseroids, n_points)``.
    """
    lib = _lib.load()
    lon_rad, sinlat = _lib.as_f64(lon_rad), _lib.as_f64(sinlat)
    coslat, radius = _lib.as_f64(coslat), _lib.as_f64(radius)
    n = lon_rad.size
    result = np.zeros(n)
    stats = np.zeros(_lib.NSTATS, dtype=np.int64)
    counts = None
    counts_ptr = None
    if leaf_counts:
        counts = np.zeros((prepared.n_tesseroids, n), dtype=np.int32)
        counts_ptr = _lib.ptr_i32(counts)
    if devices is not None:
        ids = (C.c_int * len(devices))(*[int(d) for d in devices])
        ndev = len(devices)
    else:
        ids = None
        ndev = int(njobs)
        avail = lib.tvd_device_count()
        if avail > 0 and ndev > avail:
            ndev = avail        # njobs beyond the GPU count cannot add parallelism
    rc = lib.tvd_forward(_lib.FIELDS[field], prepared.handle, n, _lib.ptr_f64(lon_rad),
                         _lib.ptr_f64(sinlat), _lib.ptr_f64(coslat), _lib.ptr_f64(radius),
                         float(ratio), int(STACK_SIZE), ndev, ids, _lib.ptr_f64(result),
                         _lib.ptr_i64(stats), counts_ptr)
    _lib.check(rc, "tvd_forward")
    if stats[3] != 0:
        warnings.warn(_WARNING_MSG, RuntimeWarning)     # tesseroid.py:227-228,232-233
    if return_stats or leaf_counts:
        out = (result, dict(zip(_lib.STAT_NAMES, (int(s) for s in stats))))
        if leaf_counts:
            out = out + (counts,)
        return out
    return result


def fused_supported(fields):
    """True if `fields` (names) can be evaluated in one fused far-field sweep."""
    lib = _lib.load()
    ids = (C.c_int * len(fields))(*[_lib.FIELDS[f] for f in fields])
    return bool(lib.tvd_fused_supported(len(fields), ids))


def forward_raw_fused(fields, lon_rad, sinlat, coslat, radius, prepared, ratios, njobs=1,
                      devices=None, return_stats=False):
    """
    Several fields in one far-field sweep (``tvd_forward_fused``): ``{field: array}`` in kernel
    units; each field's values agree with a separate :func:`forward_raw` call to a few ulp.
    """
    lib = _lib.load()
    lon_rad, sinlat = _lib.as_f64(lon_rad), _lib.as_f64(sinlat)
    coslat, radius = _lib.as_f64(coslat), _lib.as_f64(radius)
    n, nf = lon_rad.size, len(fields)
    results = np.zeros((nf, n))
    stats = np.zeros((nf, _lib.NSTATS), dtype=np.int64)
    ids = (C.c_int * nf)(*[_lib.FIELDS[f] for f in fields])
    rat = np.ascontiguousarray([float(r) for r in ratios], dtype=np.float64)
    if devices is not None:
        dev_ids = (C.c_int * len(devices))(*[int(d) for d in devices])
        ndev = len(devices)
    else:
        dev_ids = None
        ndev = int(njobs)
        avail = lib.tvd_device_count()
        if avail > 0 and ndev > avail:
            ndev = avail
    rc = lib.tvd_forward_fused(nf, ids, _lib.ptr_f64(rat), prepared.handle, n,
                               _lib.ptr_f64(lon_rad), _lib.ptr_f64(sinlat), _lib.ptr_f64(coslat),
                               _lib.ptr_f64(radius), int(STACK_SIZE), ndev, dev_ids,
                               _lib.ptr_f64(results), _lib.ptr_i64(stats))
    _lib.check(rc, "tvd_forward_fused")
    if np.any(stats[:, 3] != 0):
        warnings.warn(_WARNING_MSG, RuntimeWarning)
    out = {f: results[i] for i, f in enumerate(fields)}
    if return_stats:
        return out, {f: dict(zip(_lib.STAT_NAMES, (int(s) for s in stats[i])))
                     for i, f in enumerate(fields)}
    return out


_UNITS = {'potential': G, 'gx': SI2MGAL*G, 'gy': SI2MGAL*G, 'gz': SI2MGAL*G}
_DEFAULT_RATIO = {'potential': RATIO_V, 'gx': RATIO_G, 'gy': RATIO_G, 'gz': RATIO_G}
for _f in ('gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz'):
    _UNITS[_f] = SI2EOTVOS*G
    _DEFAULT_RATIO[_f] = RATIO_GG


def fields(names, lon, lat, height, model, dens=None, ratios=None, njobs=1, pool=None,
           delta=DELTA, devices=None):
    """
    Several fields of one model on the same points, e.g.
    ``fields(["potential", "gz", "gzz"], lon, lat, height, model)`` -> ``{name: array}``.

    The reference computes one field per call, each a full pass over points x tesseroids
    (tesseroid.py:377,435,493,556).  Here the model is flattened and discretised once, and field
    sets the far-field sweep is compiled for (see ``include/tvd.h``) share one sweep; other sets
    run field by field.  Values agree with the single-field functions to a few ulp; ``ratios``
    defaults to each field's own default (1 / 2.5 / 8).
    """
    names = list(names)
    if ratios is None:
        ratios = [_DEFAULT_RATIO[f] for f in names]
    lon, lat, height = np.asarray(lon), np.asarray(lat), np.asarray(height)
    for r in ratios:
        _check_input(lon, lat, height, model, r, njobs, pool)
    shape = lon.shape
    cc = _convert_coords(np.asarray(lon, dtype=np.float64).ravel(),
                         np.asarray(lat, dtype=np.float64).ravel(),
                         np.asarray(height, dtype=np.float64).ravel())
    own = not isinstance(model, PreparedModel)
    _check_prepared_args(model, dens, delta)
    warm = _warm_devices_async(njobs, devices) if own else None
    prepared = PreparedModel(model, dens=dens, delta=delta,
                             device=(devices[0] if devices else 0)) if own else model
    if warm is not None:
        warm.join()
    try:
        if len(names) > 1 and fused_supported(names):
            raw = forward_raw_fused(names, *cc, prepared, ratios, njobs=njobs, devices=devices)
        else:
            raw = {f: forward_raw(f, *cc, prepared, r, njobs=njobs, devices=devices)
                   for f, r in zip(names, ratios)}
    finally:
        if own:
            prepared.close()
    return {f: (raw[f] * _UNITS[f]).reshape(shape) for f in names}


def sensitivity(field, lon, lat, height, model, dens=None, ratio=None, delta=DELTA, device=0):
    """
    Sensitivity (Jacobian) matrix ``A`` of shape ``(n_points, n_tesseroids)``:
    ``A[p, t]`` = `field` at point ``p`` due to tesseroid ``t`` alone, so that
    ``A.sum(axis=1)`` is the forward-modelled field.  With ``dens=1`` this is the classic
    density sensitivity the reference's docstring mentions for the ``dens`` argument
    (tesseroid.py:347-349), which the reference builds one single-tesseroid call per column;
    column ``t`` here is bit-identical to that call.  Tesseroids the reference would skip
    (``None``, no density) get no column: columns follow the flattened model.
    """
    lib = _lib.load()
    if ratio is None:
        ratio = _DEFAULT_RATIO[field]
    lon, lat, height = np.asarray(lon), np.asarray(lat), np.asarray(height)
    _check_input(lon, lat, height, model, ratio, 1, None)
    cc = [_lib.as_f64(a) for a in _convert_coords(
        np.asarray(lon, dtype=np.float64).ravel(), np.asarray(lat, dtype=np.float64).ravel(),
        np.asarray(height, dtype=np.float64).ravel())]
    own = not isinstance(model, PreparedModel)
    _check_prepared_args(model, dens, delta)
    prepared = PreparedModel(model, dens=dens, delta=delta, device=device) if own else model
    try:
        n, nt = cc[0].size, prepared.n_tesseroids
        jac = np.zeros((nt, n))
        stats = np.zeros(_lib.NSTATS, dtype=np.int64)
        rc = lib.tvd_sensitivity(_lib.FIELDS[field], prepared.handle, n, *[_lib.ptr_f64(a) for a in cc],
                                 float(ratio), int(STACK_SIZE), int(device), _lib.ptr_f64(jac),
                                 _lib.ptr_i64(stats))
        _lib.check(rc, "tvd_sensitivity")
    finally:
        if own:
            prepared.close()
    if stats[3] != 0:
        warnings.warn(_WARNING_MSG, RuntimeWarning)
    jac *= _UNITS[field]
    return jac.T


def _dispatcher(field, lon, lat, height, model, **kwargs):
    """Counterpart of the reference's ``_dispatcher`` + ``_forward_model``
    (tesseroid.py:164-234): one preprocessing call and one batched forward call."""
    njobs = kwargs.get('njobs', 1)
    pool = kwargs.get('pool', None)
    dens = kwargs['dens']
    ratio = kwargs['ratio']
    delta = kwargs['delta']
    devices = kwargs.get('devices', None)
    lon, lat, height = np.asarray(lon), np.asarray(lat), np.asarray(height)
    result = _check_input(lon, lat, height, model, ratio, njobs, pool)
    shape = result.shape
    lon_r, sinlat, coslat, radius = _convert_coords(
        np.asarray(lon, dtype=np.float64).ravel(), np.asarray(lat, dtype=np.float64).ravel(),
        np.asarray(height, dtype=np.float64).ravel())
    own = not isinstance(model, PreparedModel)
    _check_prepared_args(model, dens, delta)
    warm = _warm_devices_async(njobs, devices) if own else None
    prepared = PreparedModel(model, dens=dens, delta=delta,
                             device=(devices[0] if devices else 0)) if own else model
    if warm is not None:
        warm.join()
    try:
        out = forward_raw(field, lon_r, sinlat, coslat, radius, prepared, ratio, njobs=njobs,
                          devices=devices)
    finally:
        if own:
            prepared.close()
    return out.reshape(shape)


def potential(lon, lat, height, model, dens=None, ratio=RATIO_V, njobs=1, pool=None,
              delta=DELTA, devices=None):
    """Gravitational potential (SI) of a tesseroid model.  Reference: tesseroid.py:327-382."""
    result = _dispatcher('potential', lon, lat, height, model, dens=dens, ratio=ratio,
                         njobs=njobs, pool=pool, delta=delta, devices=devices)
    result *= G
    return result


def gx(lon, lat, height, model, dens=None, ratio=RATIO_G, njobs=1, pool=None, delta=DELTA,
       devices=None):
    """North component of the gravitational attraction (mGal).  Reference: tesseroid.py:385-440."""
    result = _dispatcher('gx', lon, lat, height, model, dens=dens, ratio=ratio, njobs=njobs,
                         pool=pool, delta=delta, devices=devices)
    result *= SI2MGAL*G
    return result


def gy(lon, lat, height, model, dens=None, ratio=RATIO_G, njobs=1, pool=None, delta=DELTA,
       devices=None):
    """East component of the gravitational attraction (mGal).  Reference: tesseroid.py:443-498."""
    result = _dispatcher('gy', lon, lat, height, model, dens=dens, ratio=ratio, njobs=njobs,
                         pool=pool, delta=delta, devices=devices)
    result *= SI2MGAL*G
    return result


def gz(lon, lat, height, model, dens=None, ratio=RATIO_G, njobs=1, pool=None, delta=DELTA,
       devices=None):
    """Radial component of the gravitational attraction, z -> Down (mGal).
    Reference: tesseroid.py:501-561."""
    result = _dispatcher('gz', lon, lat, height, model, dens=dens, ratio=ratio, njobs=njobs,
                         pool=pool, delta=delta, devices=devices)
    result *= SI2MGAL*G
    return result


def _tensor(field):
    def func(lon, lat, height, model, dens=None, ratio=RATIO_GG, njobs=1, pool=None,
             delta=DELTA, devices=None):
        result = _dispatcher(field, lon, lat, height, model, dens=dens, ratio=ratio, njobs=njobs,
                             pool=pool, delta=delta, devices=devices)
        result *= SI2EOTVOS*G
        return result
    func.__name__ = field
    func.__doc__ = (
        "The {} component of the gravity gradient tensor (Eotvos), z -> up.  Listed in the "
        "reference docstring (tesseroid.py:56-65); integrand of Fatiando 0.5 "
        "(SURVEY.md section 8 a15).".format(field))
    return func


gxx, gxy, gxz = _tensor('gxx'), _tensor('gxy'), _tensor('gxz')
gyy, gyz, gzz = _tensor('gyy'), _tensor('gyz'), _tensor('gzz')


def _density_based_discretization(bounds, density, delta):
    """
    Density-based radial discretisation of one tesseroid, run by the K1 preprocessing kernel.
    Same contract as the reference (tesseroid.py:237-283): returns the list of
    ``[w, e, s, n, top, bottom]`` arrays in the reference's (FIFO) emission order.
    """
    bounds = np.asarray(bounds, dtype=np.float64)
    kind, params = as_kind_params(density)
    flat = FlatModel(bounds.reshape(1, 6), np.array([kind], dtype=np.int32),
                     np.array([params], dtype=np.float64).reshape(1, NPAR))
    prepared = PreparedModel(flat, delta=delta)
    try:
        top, bottom, _ = prepared.pieces()
    finally:
        prepared.close()
    w, e, s, n = bounds[:4]
    return [np.array([w, e, s, n, t, b]) for t, b in zip(top, bottom)]


def _split_arrays(arrays, extra_args, nparts):
    """
    Split the coordinate arrays into nparts. Add extra_args to each part.
    Same chunking as the reference (tesseroid.py:303-324); libtvd applies the identical rule
    when it shards points over GPUs.

    >>> chunks = _split_arrays([[1, 2, 3, 4, 5, 6]], ['meh'], 3)
    >>> chunks[0]
    [[1, 2], 'meh']
    >>> chunks[2]
    [[5, 6], 'meh']
    """
    size = len(arrays[0])
    n = size//nparts
    strides = [(i*n, (i + 1)*n) for i in range(nparts - 1)]
    strides.append((strides[-1][-1] if strides else 0, size))
    chunks = [[x[low:high] for x in arrays] + extra_args
              for low, high in strides]
    return chunks
