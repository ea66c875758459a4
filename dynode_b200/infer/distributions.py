"""Prior / observation distributions and bijectors in torch (float64).

The reference takes these from numpyro (`numpyro.distributions`, reference pyproject.toml:17), which is
absent from this image.  Only what DynODE's inference path touches is provided, with numpyro's
semantics: `log_prob`, `sample`, `support`, and `biject_to(support)` mapping the unconstrained space
NUTS/SVI work in onto the support (interval -> sigmoid then affine, positive -> exp), including the
log-abs-det-Jacobian that enters the log-density.

Every function here is plain elementwise torch, so it runs under `torch.vmap` (the per-draw model is
vmapped over chains) and on the device next to the ODE kernels.
"""

from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

_F64 = torch.float64


def _t(x, like: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Numbers become 0-dim HOST tensors: they mix with device tensors as scalars, and no host-to-device
    copy is issued (such a copy would break CUDA-graph capture of the model evaluation)."""
    if isinstance(x, torch.Tensor):
        return x if x.dtype == _F64 else x.to(_F64)
    return torch.as_tensor(x, dtype=_F64)


# ------------------------------------------------------------------------------ constraints
class Constraint:
    def __repr__(self):
        return type(self).__name__.lstrip("_")


class _Real(Constraint):
    pass


class _Positive(Constraint):
    pass


class _UnitInterval(Constraint):
    lower_bound, upper_bound = 0.0, 1.0


class _NonnegativeInteger(Constraint):
    pass


class _Interval(Constraint):
    def __init__(self, lower_bound, upper_bound):
        self.lower_bound, self.upper_bound = lower_bound, upper_bound

    def __repr__(self):
        return f"Interval({self.lower_bound}, {self.upper_bound})"


class _GreaterThan(Constraint):
    def __init__(self, lower_bound):
        self.lower_bound = lower_bound


class _LessThan(Constraint):
    def __init__(self, upper_bound):
        self.upper_bound = upper_bound


class constraints:
    real = _Real()
    positive = _Positive()
    unit_interval = _UnitInterval()
    nonnegative_integer = _NonnegativeInteger()
    interval = _Interval
    greater_than = _GreaterThan
    less_than = _LessThan


# ------------------------------------------------------------------------------ transforms
class Transform:
    def __call__(self, x):
        raise NotImplementedError

    def inv(self, y):
        raise NotImplementedError

    def log_abs_det_jacobian(self, x, y):
        raise NotImplementedError

    def codomain(self, domain: Constraint) -> Constraint:
        return constraints.real


class IdentityTransform(Transform):
    def __call__(self, x):
        return x

    def inv(self, y):
        return y

    def log_abs_det_jacobian(self, x, y):
        return torch.zeros_like(x)

    def codomain(self, domain):
        return domain


class ExpTransform(Transform):
    def __call__(self, x):
        return torch.exp(x)

    def inv(self, y):
        return torch.log(y)

    def log_abs_det_jacobian(self, x, y):
        return x

    def codomain(self, domain):
        return constraints.positive


class SigmoidTransform(Transform):
    def __call__(self, x):
        return torch.sigmoid(x)

    def inv(self, y):
        return torch.log(y) - torch.log1p(-y)

    def log_abs_det_jacobian(self, x, y):
        return -torch.nn.functional.softplus(x) - torch.nn.functional.softplus(-x)

    def codomain(self, domain):
        return constraints.unit_interval


class AffineTransform(Transform):
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def __call__(self, x):
        return self.loc + self.scale * x

    def inv(self, y):
        return (y - self.loc) / self.scale

    def log_abs_det_jacobian(self, x, y):
        return torch.log(torch.abs(_t(self.scale, x))) + torch.zeros_like(x)

    def codomain(self, domain):
        if isinstance(domain, _Real):
            return domain
        lo = getattr(domain, "lower_bound", None)
        hi = getattr(domain, "upper_bound", None)
        if isinstance(domain, _Positive):
            lo = 0.0
        a = None if lo is None else self.loc + self.scale * lo
        b = None if hi is None else self.loc + self.scale * hi
        if float(torch.as_tensor(self.scale).reshape(-1)[0]) < 0:
            a, b = b, a
        if a is not None and b is not None:
            return constraints.interval(a, b)
        return constraints.greater_than(a) if a is not None else constraints.less_than(b)


class ComposeTransform(Transform):
    def __init__(self, parts: Sequence[Transform]):
        self.parts = list(parts)

    def __call__(self, x):
        for p in self.parts:
            x = p(x)
        return x

    def inv(self, y):
        for p in reversed(self.parts):
            y = p.inv(y)
        return y

    def log_abs_det_jacobian(self, x, y):
        total = 0.0
        for p in self.parts:
            nxt = p(x)
            total = total + p.log_abs_det_jacobian(x, nxt)
            x = nxt
        return total


class transforms:
    AffineTransform = AffineTransform
    ExpTransform = ExpTransform
    SigmoidTransform = SigmoidTransform
    IdentityTransform = IdentityTransform
    ComposeTransform = ComposeTransform


def biject_to(support: Constraint) -> Transform:
    """Unconstrained reals -> `support`, with numpyro's choice of bijector per constraint."""
    if isinstance(support, _Real):
        return IdentityTransform()
    if isinstance(support, _Positive):
        return ExpTransform()
    if isinstance(support, _UnitInterval):
        return SigmoidTransform()
    if isinstance(support, _Interval):
        return ComposeTransform([SigmoidTransform(),
                                 AffineTransform(support.lower_bound, support.upper_bound - support.lower_bound)])
    if isinstance(support, _GreaterThan):
        return ComposeTransform([ExpTransform(), AffineTransform(support.lower_bound, 1.0)])
    if isinstance(support, _LessThan):
        return ComposeTransform([ExpTransform(), AffineTransform(support.upper_bound, -1.0)])
    raise NotImplementedError(f"no bijector for support {support!r}")


# ---- fused bijector + log-Jacobian on the device (include/dynode_b200_ppl.h) ---------------------------
BIJ_INTERVAL, BIJ_GREATER_THAN, BIJ_LESS_THAN, BIJ_REAL = 0, 1, 2, 3


class _FusedBijector(torch.autograd.Function):
    """(z) -> (x, log|dx/dz|) for biject_to(interval / greater_than / less_than) in ONE elementwise kernel,
    its vector-Jacobian product in another: `dynode_bijector_f64` / `dynode_bijector_vjp_f64`.  The same map
    written with tensor operations is ~13 launches forward and ~12 backward per site, which at a few thousand
    NUTS chains is most of a round."""

    @staticmethod
    def forward(z, kind: int, a: float, b: float):
        import ctypes

        from .. import _lib
        zc = z.contiguous()
        x, ladj = torch.empty_like(zc), torch.empty_like(zc)
        _lib.check(_lib.load().dynode_bijector_f64(int(kind), zc.numel(), zc.data_ptr(), float(a), float(b),
                                                   x.data_ptr(), ladj.data_ptr(),
                                                   ctypes.c_void_p(_lib.current_stream_ptr())))
        return x, ladj

    @staticmethod
    def setup_context(ctx, inputs, output):
        z, kind, a, b = inputs
        ctx.kind, ctx.b = int(kind), float(b)
        ctx.save_for_backward(z)

    @staticmethod
    def backward(ctx, gx, gl):
        import ctypes

        from .. import _lib
        (z,) = ctx.saved_tensors
        zc, gxc, glc = z.contiguous(), gx.contiguous(), gl.contiguous()
        gz = torch.empty_like(zc)
        _lib.check(_lib.load().dynode_bijector_vjp_f64(ctx.kind, zc.numel(), zc.data_ptr(), ctx.b, gxc.data_ptr(),
                                                       glc.data_ptr(), gz.data_ptr(),
                                                       ctypes.c_void_p(_lib.current_stream_ptr())))
        return gz, None, None, None

    @staticmethod
    def vmap(info, in_dims, z, kind, a, b):  # elementwise: the batch axis is just one more axis
        x, ladj = _FusedBijector.apply(z.movedim(in_dims[0], 0), kind, a, b)
        return (x, ladj), (0, 0)


def _host_float(v) -> Optional[float]:
    """A bound as a python float when it is a number or a 0-dim HOST tensor (no device sync), else None."""
    if isinstance(v, (int, float)):
        return float(v)
    if isinstance(v, torch.Tensor) and v.device.type == "cpu" and v.numel() == 1 and not v.requires_grad \
            and not torch._C._functorch.is_batchedtensor(v):
        return float(v)
    return None


FAM_NORMAL, FAM_UNIFORM, FAM_BETA, FAM_GAMMA, FAM_LOGNORMAL, FAM_HALFNORMAL, FAM_EXPONENTIAL = range(7)


def _bijector_spec(support: Constraint):
    """(kind, a, b) of the fused bijector for a support with constant bounds, else None."""
    if isinstance(support, _Real):
        return BIJ_REAL, 0.0, 1.0
    if isinstance(support, _UnitInterval):
        return BIJ_INTERVAL, 0.0, 1.0
    if isinstance(support, _Interval):
        lo, hi = _host_float(support.lower_bound), _host_float(support.upper_bound)
        return None if lo is None or hi is None else (BIJ_INTERVAL, lo, hi - lo)
    if isinstance(support, _Positive):
        return BIJ_GREATER_THAN, 0.0, 1.0
    if isinstance(support, _GreaterThan):
        lo = _host_float(support.lower_bound)
        return None if lo is None else (BIJ_GREATER_THAN, lo, 1.0)
    if isinstance(support, _LessThan):
        hi = _host_float(support.upper_bound)
        return None if hi is None else (BIJ_LESS_THAN, hi, 1.0)
    return None


def _family_spec(fn):
    """(family, p0, p1, c, aff_loc, aff_scale) for a prior whose parameters are constants (numbers / 0-dim host
    tensors), with c = every x-independent term of its log_prob; None if the fused site kernel cannot state it."""
    hf = _host_float
    if isinstance(fn, TransformedDistribution) and isinstance(fn.transform, AffineTransform):
        base, loc, sc = _family_spec(fn.base_dist), hf(fn.transform.loc), hf(fn.transform.scale)
        if base is None or loc is None or sc is None or sc == 0.0 or base[4:] != (0.0, 1.0):
            return None
        return base[0], base[1], base[2], base[3] - math.log(abs(sc)), loc, sc
    vals = [hf(v) for v in fn._params()] if type(fn) in (Normal, LogNormal, HalfNormal, Exponential, Uniform,
                                                         Gamma, Beta, TruncatedNormal) else [None]
    if any(v is None for v in vals):
        return None
    half_log_2pi = 0.5 * math.log(2.0 * math.pi)
    if type(fn) is Normal:
        return FAM_NORMAL, vals[0], vals[1], -math.log(vals[1]) - half_log_2pi, 0.0, 1.0
    if type(fn) is TruncatedNormal:
        lo = None if fn.low is None else hf(fn.low)
        hi = None if fn.high is None else hf(fn.high)
        if (fn.low is not None and lo is None) or (fn.high is not None and hi is None):
            return None
        a, b = fn._cdf_bounds()  # host scalars
        return FAM_NORMAL, vals[0], vals[1], -math.log(vals[1]) - half_log_2pi - math.log(float(b - a)), 0.0, 1.0
    if type(fn) is LogNormal:
        return FAM_LOGNORMAL, vals[0], vals[1], -math.log(vals[1]) - half_log_2pi, 0.0, 1.0
    if type(fn) is HalfNormal:
        return FAM_HALFNORMAL, vals[0], 0.0, -math.log(vals[0]) + 0.5 * math.log(2.0 / math.pi), 0.0, 1.0
    if type(fn) is Exponential:
        return FAM_EXPONENTIAL, vals[0], 0.0, math.log(vals[0]), 0.0, 1.0
    if type(fn) is Uniform:
        return FAM_UNIFORM, 0.0, 0.0, -math.log(vals[1] - vals[0]), 0.0, 1.0
    if type(fn) is Gamma:
        return FAM_GAMMA, vals[0], vals[1], vals[0] * math.log(vals[1]) - math.lgamma(vals[0]), 0.0, 1.0
    if type(fn) is Beta:
        a, b = vals
        return FAM_BETA, a, b, -(math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)), 0.0, 1.0
    return None


class _FusedSite(torch.autograd.Function):
    """z -> (x, log|dx/dz| + log p(x)) for a latent site whose prior has constant parameters: ONE elementwise
    kernel forward and one backward (`dynode_site_logdensity_f64` / `_vjp_f64`) in place of the ~25 launches the
    composed transforms and log_prob take each way.  A 1-d `z` is read through its stride (a site is usually a
    column of the sampler's [chains, D] array), so no gather copy is made either."""

    @staticmethod
    def _strided(z):
        if z.dim() == 1 and z.numel() > 0 and z.stride(0) > 0:
            return z, int(z.stride(0))
        return z.contiguous(), 1

    @staticmethod
    def forward(z, spec):
        import ctypes

        from .. import _lib
        zc, zs = _FusedSite._strided(z)
        x = torch.empty(z.shape, dtype=z.dtype, device=z.device)
        lp = torch.empty(z.shape, dtype=z.dtype, device=z.device)
        sd = _lib.SiteDesc(*spec)
        _lib.check(_lib.load().dynode_site_logdensity_f64(ctypes.byref(sd), z.numel(), zc.data_ptr(), zs, x.data_ptr(),
                                                          lp.data_ptr(), ctypes.c_void_p(_lib.current_stream_ptr())))
        return x, lp

    @staticmethod
    def setup_context(ctx, inputs, output):
        z, spec = inputs
        ctx.spec = spec
        ctx.save_for_backward(z)

    @staticmethod
    def backward(ctx, gx, glp):
        import ctypes

        from .. import _lib
        (z,) = ctx.saved_tensors
        zc, zs = _FusedSite._strided(z)
        gxc, glc = gx.contiguous(), glp.contiguous()
        gz = torch.empty(z.shape, dtype=z.dtype, device=z.device)
        sd = _lib.SiteDesc(*ctx.spec)
        _lib.check(_lib.load().dynode_site_logdensity_vjp_f64(ctypes.byref(sd), z.numel(), zc.data_ptr(), zs,
                                                              gxc.data_ptr(), glc.data_ptr(), gz.data_ptr(),
                                                              ctypes.c_void_p(_lib.current_stream_ptr())))
        return gz, None

    @staticmethod
    def vmap(info, in_dims, z, spec):
        x, lp = _FusedSite.apply(z.movedim(in_dims[0], 0), spec)
        return (x, lp), (0, 0)


def fused_site(fn, z: torch.Tensor):
    """(x, log|dx/dz| + fn.log_prob(x)) from one kernel when `z` lives on a CUDA device and the prior `fn` has
    constant parameters and a supported family; None otherwise (the caller composes transforms + log_prob)."""
    if not (z.is_cuda and z.dtype == _F64):
        return None
    spec = getattr(fn, "_fused_site_spec", False)
    if spec is False:
        bij, fam = _bijector_spec(fn.support), _family_spec(fn)
        spec = None if bij is None or fam is None else (bij[0], fam[0], bij[1], bij[2], fam[1], fam[2], fam[3],
                                                        fam[4], fam[5])
        try:
            fn._fused_site_spec = spec  # the prior objects of a config are reused by every model call
        except AttributeError:
            pass
    return None if spec is None else _FusedSite.apply(z, spec)


def constrain_with_ladj(support: Constraint, z: torch.Tensor):
    """x = biject_to(support)(z) and log|dx/dz| (elementwise).  Latent sites on a CUDA device whose support has
    constant bounds take the fused kernels; everything else the composed transforms."""
    if z.is_cuda and z.dtype == _F64:
        spec = _bijector_spec(support)
        if spec is not None and spec[0] != BIJ_REAL:
            return _FusedBijector.apply(z, *spec)
    t = biject_to(support)
    x = t(z)
    return x, t.log_abs_det_jacobian(z, x)


# ------------------------------------------------------------------------------ distributions
def _std_normal_cdf(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _std_normal_icdf(p):
    return math.sqrt(2.0) * torch.erfinv(2.0 * p - 1.0)


class Distribution:
    """Batch shape = broadcast shape of the parameters; event shape is always () here."""

    support: Constraint = constraints.real
    is_discrete = False

    def _params(self) -> Sequence[torch.Tensor]:
        raise NotImplementedError

    @property
    def batch_shape(self):
        return torch.broadcast_shapes(*[p.shape for p in self._params()])

    def shape(self, sample_shape=()):
        return tuple(sample_shape) + tuple(self.batch_shape)

    def _device(self):
        return self._params()[0].device

    def _uniform(self, generator, sample_shape):
        return torch.rand(self.shape(sample_shape), dtype=_F64, device=self._device(), generator=generator)

    def _normal(self, generator, sample_shape):
        return torch.randn(self.shape(sample_shape), dtype=_F64, device=self._device(), generator=generator)

    def _generator(self, key):
        """numpyro passes a PRNGKey as `key`; here it may also be a torch.Generator or None (global RNG)."""
        if key is None or isinstance(key, torch.Generator):
            return key
        return key.generator(self._device())

    def sample(self, key=None, sample_shape=()):
        raise NotImplementedError

    def log_prob(self, value):
        raise NotImplementedError

    def to(self, device):
        """Copy of the distribution with its parameters on `device`."""
        import copy
        new = copy.copy(self)
        for k, v in vars(self).items():
            if isinstance(v, torch.Tensor):
                setattr(new, k, v.to(device))
            elif isinstance(v, Distribution):
                setattr(new, k, v.to(device))
        return new


class Normal(Distribution):
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = _t(loc), _t(scale)

    def _params(self):
        return (self.loc, self.scale)

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        return self.loc + self.scale * self._normal(generator, sample_shape)

    def log_prob(self, value):
        z = (value - self.loc) / self.scale
        return -0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi)


class LogNormal(Distribution):
    support = constraints.positive

    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = _t(loc), _t(scale)

    def _params(self):
        return (self.loc, self.scale)

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        return torch.exp(self.loc + self.scale * self._normal(generator, sample_shape))

    def log_prob(self, value):
        lv = torch.log(value)
        z = (lv - self.loc) / self.scale
        return -0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi) - lv


class HalfNormal(Distribution):
    support = constraints.positive

    def __init__(self, scale=1.0):
        self.scale = _t(scale)

    def _params(self):
        return (self.scale,)

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        return torch.abs(self.scale * self._normal(generator, sample_shape))

    def log_prob(self, value):
        z = value / self.scale
        return -0.5 * z * z - torch.log(self.scale) + 0.5 * math.log(2.0 / math.pi)


class Exponential(Distribution):
    support = constraints.positive

    def __init__(self, rate=1.0):
        self.rate = _t(rate)

    def _params(self):
        return (self.rate,)

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        return -torch.log1p(-self._uniform(generator, sample_shape)) / self.rate

    def log_prob(self, value):
        return torch.log(self.rate) - self.rate * value


class Uniform(Distribution):
    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = _t(low), _t(high)
        self.support = constraints.interval(self.low, self.high)

    def _params(self):
        return (self.low, self.high)

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        return self.low + (self.high - self.low) * self._uniform(generator, sample_shape)

    def log_prob(self, value):
        inside = (value >= self.low) & (value <= self.high)
        lp = -torch.log(self.high - self.low) + torch.zeros_like(value)
        return torch.where(inside, lp, torch.full_like(lp, -math.inf))


class Gamma(Distribution):
    support = constraints.positive

    def __init__(self, concentration, rate=1.0):
        self.concentration, self.rate = _t(concentration), _t(rate)

    def _params(self):
        return (self.concentration, self.rate)

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        a = self.concentration.expand(self.shape(sample_shape)).contiguous()
        g = torch._standard_gamma(a, generator=generator) if generator is not None else torch._standard_gamma(a)
        return g / self.rate

    def log_prob(self, value):
        a, b = self.concentration, self.rate
        return a * torch.log(b) + (a - 1.0) * torch.log(value) - b * value - torch.lgamma(a)


class Beta(Distribution):
    support = constraints.unit_interval

    def __init__(self, concentration1, concentration0):
        self.concentration1, self.concentration0 = _t(concentration1), _t(concentration0)

    def _params(self):
        return (self.concentration1, self.concentration0)

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        shp = self.shape(sample_shape)
        kw = {} if generator is None else {"generator": generator}
        ga = torch._standard_gamma(self.concentration1.expand(shp).contiguous(), **kw)
        gb = torch._standard_gamma(self.concentration0.expand(shp).contiguous(), **kw)
        return ga / (ga + gb)

    def log_prob(self, value):
        a, b = self.concentration1, self.concentration0
        log_beta = torch.lgamma(a) + torch.lgamma(b) - torch.lgamma(a + b)
        return torch.xlogy(a - 1.0, value) + torch.xlogy(b - 1.0, 1.0 - value) - log_beta


class TruncatedNormal(Distribution):
    """Normal(loc, scale) restricted to [low, high] (either bound may be None)."""

    def __init__(self, loc=0.0, scale=1.0, *, low=None, high=None):
        self.loc, self.scale = _t(loc), _t(scale)
        self.low = None if low is None else _t(low)
        self.high = None if high is None else _t(high)
        if low is not None and high is not None:
            self.support = constraints.interval(self.low, self.high)
        elif low is not None:
            self.support = constraints.greater_than(self.low)
        elif high is not None:
            self.support = constraints.less_than(self.high)

    def _params(self):
        return (self.loc, self.scale)

    @property
    def base_dist(self):
        return Normal(self.loc, self.scale)

    def _cdf_bounds(self):
        zero, one = torch.zeros_like(self.loc), torch.ones_like(self.loc)
        a = zero if self.low is None else _std_normal_cdf((self.low - self.loc) / self.scale)
        b = one if self.high is None else _std_normal_cdf((self.high - self.loc) / self.scale)
        return a, b

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        a, b = self._cdf_bounds()
        u = self._uniform(generator, sample_shape)
        p = (a + u * (b - a)).clamp(1e-300, 1.0 - 1e-16)
        x = self.loc + self.scale * _std_normal_icdf(p)
        if self.low is not None:
            x = torch.maximum(x, self.low)
        if self.high is not None:
            x = torch.minimum(x, self.high)
        return x

    def log_prob(self, value):
        a, b = self._cdf_bounds()
        z = (value - self.loc) / self.scale
        return -0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi) - torch.log(b - a)


class Poisson(Distribution):
    support = constraints.nonnegative_integer
    is_discrete = True

    def __init__(self, rate):
        self.rate = _t(rate)

    def _params(self):
        return (self.rate,)

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        r = self.rate.expand(self.shape(sample_shape))
        return torch.poisson(r, generator=generator) if generator is not None else torch.poisson(r)

    def log_prob(self, value):
        # numpyro: value*log(rate) - lgamma(value + 1) - rate; non-integer observations are accepted
        # (the reference's synthetic incidence is un-noised, examples/sir_infer_parameters.py:86-108)
        return torch.xlogy(value, self.rate) - torch.lgamma(value + 1.0) - self.rate


class TransformedDistribution(Distribution):
    def __init__(self, base_distribution: Distribution, transforms_):
        self.base_dist = base_distribution
        parts = list(transforms_) if isinstance(transforms_, (list, tuple)) else [transforms_]
        self.transform = parts[0] if len(parts) == 1 else ComposeTransform(parts)
        sup = base_distribution.support
        for p in parts:
            sup = p.codomain(sup)
        self.support = sup

    def _params(self):
        return self.base_dist._params()

    def sample(self, key=None, sample_shape=()):
        generator = self._generator(key)
        return self.transform(self.base_dist.sample(generator, sample_shape))

    def log_prob(self, value):
        x = self.transform.inv(value)
        return self.base_dist.log_prob(x) - self.transform.log_abs_det_jacobian(x, value)


__all__ = ["Distribution", "Normal", "LogNormal", "HalfNormal", "Exponential", "Uniform", "Gamma", "Beta",
           "TruncatedNormal", "Poisson", "TransformedDistribution", "constraints", "transforms", "biject_to",
           "Transform", "AffineTransform", "ExpTransform", "SigmoidTransform", "IdentityTransform",
           "ComposeTransform", "Constraint"]
