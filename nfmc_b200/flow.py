"""Host-side mirror of the ``torchflows`` surface nfmc uses, backed by the sm_100a kernels.

``Flow`` / ``RealNVP`` are parameter containers (``nn.Module``; same ``state_dict`` keys as the oracle's
restatement) whose operators -- ``log_prob``, ``sample``, ``bijection.forward`` / ``inverse`` -- run on the
GPU through ``libnfmc_b200.so``.  The call sites they serve in the reference:
``/root/reference/nfmc/algorithms/sampling/base.py:26`` (default flow), ``nfmc/util.py:280-281``
(``flow='realnvp'``), ``nfmc/jump.py:205,218``, ``nfmc/imh.py:214,221``, ``nfmc/neutra.py:60``.

Arithmetic convention (see ``oracle/realnvp_ref.py`` -- parity against torchflows itself is unpinned because
that package is absent): affine pair ``(u_a, u_b)`` -> ``alpha = exp(log(1-m) + u_a/2) + m``, ``m = 1e-3``,
``beta = u_b/2``; layers ``[Affine] + Lc x [Reverse, Coupling, ActNorm] + [Affine, ActNorm]``.

``pack_realnvp`` folds the reverse permutations into the parameter order so the kernels never permute the
chain state ("flow blob").
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _native as N

MIN_SCALE = 1e-3
_LOG_ONE_MINUS_M = math.log(1.0 - MIN_SCALE)


def _affine(u_a: torch.Tensor, u_b: torch.Tensor):
    alpha = torch.exp(_LOG_ONE_MINUS_M + u_a / 2) + MIN_SCALE
    return alpha, torch.log(alpha), u_b / 2


class _Layer(nn.Module):
    def __init__(self, event_shape):
        super().__init__()
        self.event_shape = tuple(int(s) for s in event_shape)
        self.n_dim = int(math.prod(self.event_shape))


class ElementwiseAffine(_Layer):
    def __init__(self, event_shape):
        super().__init__(event_shape)
        self.value = nn.Parameter(torch.zeros(self.n_dim, 2))


class ActNorm(ElementwiseAffine):
    def __init__(self, event_shape):
        super().__init__(event_shape)
        self.register_buffer("initialised", torch.tensor(False))


class ReversePermutation(_Layer):
    pass


def default_hidden(n_source: int) -> int:
    return max(int(3 * math.log10(n_source)), 4)


class AffineCoupling(_Layer):
    def __init__(self, event_shape, conditioner_kwargs: Optional[dict] = None, **_ignored):
        super().__init__(event_shape)
        ck = dict(conditioner_kwargs or {})
        self.n_source = self.n_dim // 2
        self.n_target = self.n_dim - self.n_source
        self.n_linear = int(ck.get("n_layers", 2))
        hidden = ck.get("n_hidden", None)
        self.n_hidden = int(default_hidden(self.n_source) if hidden is None else hidden)
        if self.n_linear < 1:
            raise ValueError("conditioner needs at least one linear layer")
        mods = []
        if self.n_linear == 1:
            mods.append(nn.Linear(self.n_source, 2 * self.n_target))
        else:
            mods += [nn.Linear(self.n_source, self.n_hidden), nn.Tanh()]
            for _ in range(self.n_linear - 2):
                mods += [nn.Linear(self.n_hidden, self.n_hidden), nn.Tanh()]
            mods.append(nn.Linear(self.n_hidden, 2 * self.n_target))
        self.net = nn.Sequential(*mods)
        with torch.no_grad():  # identity map at initialisation
            self.net[-1].weight.zero_()
            self.net[-1].bias.zero_()

    def linears(self):
        return [m for m in self.net if isinstance(m, nn.Linear)]


class RealNVP(_Layer):
    """RealNVP bijection.  ``forward``: data -> latent, ``inverse``: latent -> data; both return ``(y, log_det)``."""

    def __init__(self, event_shape, n_layers: int = 2, edge_list=None, conditioner_dtype: str = "auto", **kwargs):
        """``conditioner_dtype``: ``'fp32'`` = CUDA-core kernels (rtol 1e-4); ``'bf16'`` = tcgen05 tensor-core kernel
        (bf16 operands, fp32 accumulate, rtol 1e-2; needs 2 linear layers, hidden a multiple of 16 in [16, 256],
        even d <= 128); ``'auto'`` = bf16 whenever the shape is eligible."""
        if isinstance(event_shape, int):
            event_shape = (event_shape,)
        super().__init__(event_shape)
        if conditioner_dtype not in ("auto", "fp32", "bf16"):
            raise ValueError("conditioner_dtype must be 'auto', 'fp32' or 'bf16'")
        self.conditioner_dtype = conditioner_dtype
        if edge_list is not None:
            raise NotImplementedError("edge_list couplings are outside the accelerated hot path")
        if self.n_dim < 2 or self.n_dim > N.MAX_DIM:
            raise ValueError(f"RealNVP event size must be in [2, {N.MAX_DIM}]")
        layers = [ElementwiseAffine(event_shape)]
        for _ in range(int(n_layers)):
            layers += [ReversePermutation(event_shape), AffineCoupling(event_shape, **kwargs), ActNorm(event_shape)]
        layers += [ElementwiseAffine(event_shape), ActNorm(event_shape)]
        self.layers = nn.ModuleList(layers)
        self.n_coupling = int(n_layers)
        self._packed = {}  # device -> (version key, blob tensor)
        self._packed_tc = {}
        if conditioner_dtype == "bf16" and not tc_supported(self.n_dim, *self.conditioner_shape()):
            raise ValueError("conditioner_dtype='bf16' needs 2 linear layers, hidden <= 256, even d <= 128")

    # -- structure --------------------------------------------------------------------------------------------
    def couplings(self):
        return [m for m in self.layers if isinstance(m, AffineCoupling)]

    def conditioner_shape(self) -> Tuple[int, int]:
        cs = self.couplings()
        if not cs:
            return 2, 4
        return cs[0].n_linear, cs[0].n_hidden

    def _version_key(self):
        return tuple(p._version for p in self.parameters()) + tuple(id(p) for p in self.parameters())

    def blob(self, device: torch.device) -> torch.Tensor:
        key = str(device)
        ver = self._version_key()
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            self._packed[key] = (ver, self._pack_on(device))
        return self._packed[key][1]

    def _pack_on(self, device: torch.device) -> torch.Tensor:
        """Packed blob on ``device``.  Parameters that already live on that GPU are packed there by ``nfmc_flow_pack``
        (one launch; this is the path after every flow refit); otherwise on the host by ``pack_realnvp``."""
        params = list(self.parameters())
        M, H = self.conditioner_shape()
        same = all((c.n_linear, c.n_hidden) == (M, H) for c in self.couplings())
        device = torch.device(device)
        idx = device.index if device.index is not None else (torch.cuda.current_device() if device.type == 'cuda' else None)
        on_dev = device.type == 'cuda' and all(p.is_cuda and p.device.index == idx for p in params)
        if params and same and on_dev and M == 2 and H <= SMALL_H:
            P = N.lib().nfmc_flow_param_count(self.n_dim, self.n_coupling, M, H)
            theta = torch.cat([p.detach().reshape(-1) for p in params]).to(torch.float32).contiguous()
            if P == theta.numel():
                blob = torch.empty(blob_floats(self.n_dim, self.n_coupling, M, H), device=device, dtype=torch.float32)
                N.check(N.lib().nfmc_flow_pack(self.n_dim, self.n_coupling, M, H, N.ptr(theta), N.ptr(blob), N.stream_ptr(device)))
                return blob
        return pack_realnvp(self).to(device)

    def uses_tensor_cores(self) -> bool:
        shape = (self.n_dim, *self.conditioner_shape())
        if self.conditioner_dtype == "bf16":
            return tc_supported(*shape)
        return self.conditioner_dtype == "auto" and tc_eligible(*shape)

    def uses_row_tile_pass(self) -> bool:
        """Conditioner shapes outside the register-resident path (M = 2, H <= 8) and outside the tensor-core path -- deep
        conditioners (n_layers != 2, e.g. the reference's ``n_layers=5, n_hidden=100``), odd d or d > 128 with H > 8, or any
        wide conditioner with ``conditioner_dtype='fp32'`` -- take the row-tile fp32 pass of csrc/train_wide.cu
        (``nfmc_flow_wide_pass`` / ``nfmc_flow_wide_log_prob`` / ``nfmc_jump_step_wide``) for forward / inverse / log_prob and
        the NF jump / IMH step: register-blocked weight reuse over a tile of rows instead of one weight load per chain and
        multiply (3-6x the per-chain generic path, ``tools/bench_deep.py``)."""
        return self.row_tile_supported() and not self.uses_tensor_cores()

    def row_tile_supported(self) -> bool:
        """Shape test of the row-tile fp32 kernels: any conditioner outside the register-resident path whose couplings share
        one shape.  (NeuTra also takes them for a tensor-core flow whose shape the tensor-core NeuTra kernels refuse, e.g.
        d % 4 != 0: value and gradient of the latent potential then both come from the fp32 pass.)"""
        M, H = self.conditioner_shape()
        if (M == 2 and H <= SMALL_H) or not self.couplings():
            return False
        if os.environ.get("NFMC_B200_NO_ROW_TILE") == "1":      # tests / A-B runs: keep the per-chain generic conditioner
            return False
        same = all((c.n_linear, c.n_hidden) == (M, H) for c in self.couplings())
        return same and N.lib().nfmc_flow_wide_param_count(self.n_dim, self.n_coupling, M, H) > 0

    def theta_descriptor(self, device: torch.device, transposed: bool = True):
        """Descriptor of the row-tile path: the module-order parameter vector (cached per device and parameter version).
        ``transposed`` (what the pass / log_prob / sample / jump entry points take with their ``transposed`` flag set): every
        linear's weight stored ``[in][out]`` at the same offset, so that the forward GEMMs, whose lanes run over output units,
        read it coalesced; ``transposed=False`` is the modules' own ``[out][in]`` order (the backward sweep and the trainer)."""
        key = ("thetaT:" if transposed else "theta:") + str(device)
        ver = self._version_key()
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            lin = {id(m.weight) for c in self.couplings() for m in c.linears()} if transposed else set()
            parts = [(p.detach().t() if id(p) in lin else p.detach()).reshape(-1) for p in self.parameters()]
            self._packed[key] = (ver, torch.cat(parts).to(device, torch.float32).contiguous())
        theta = self._packed[key][1]
        M, H = self.conditioner_shape()
        return N.RealNVPDesc(self.n_dim, self.n_coupling, M, H, theta.data_ptr(), theta.numel()), theta

    def tc_descriptor(self, device: torch.device):
        key = str(device)
        ver = self._version_key()
        hit = self._packed_tc.get(key)
        if hit is None or hit[0] != ver:
            self._packed_tc[key] = (ver, pack_realnvp_tc(self).to(device))
        blob = self._packed_tc[key][1]
        M, H = self.conditioner_shape()
        return N.RealNVPTcDesc(self.n_dim, self.n_coupling, _tc_hidden(H), 0, blob.data_ptr(), blob.numel()), blob

    def tc_transposed(self, device: torch.device) -> torch.Tensor:
        """Transposed weight images for the tensor-core NeuTra kernel (``pack_realnvp_tc_transposed``), cached per device."""
        key = str(device) + "/T"
        ver = self._version_key()
        hit = self._packed_tc.get(key)
        if hit is None or hit[0] != ver:
            self._packed_tc[key] = (ver, pack_realnvp_tc_transposed(self).to(device))
        return self._packed_tc[key][1]

    #: 'auto' sends a DEFAULT (narrow, H <= 8) conditioner to the tensor-core NeuTra kernels from this many chains on: there the
    #: tcgen05 path is 1.5x the fp32 CUDA-core kernel (C3, 262 144 chains: 3.76e7 vs 2.51e7 chain-steps/s), below it the run is
    #: launch-bound and the single fused fp32 launch wins
    NEUTRA_AUTO_MIN_CHAINS = 32768

    def uses_tensor_cores_for_neutra(self, n_chains: Optional[int] = None) -> bool:
        """Whether NeuTra-HMC takes the tensor-core kernels (csrc/tc_neutra.cu) for this flow.  ``conditioner_dtype='bf16'``:
        whenever the shape is eligible; ``'fp32'``: never; ``'auto'``: wide conditioners (H > 8) always, default ones only for
        ``n_chains >= NEUTRA_AUTO_MIN_CHAINS`` (the north star's bf16-conditioner tolerance, rtol 1e-2, applies on that path)."""
        shape = (self.n_dim, *self.conditioner_shape())
        if not neutra_tc_supported(*shape, self.n_coupling):
            return False
        if self.conditioner_dtype == "bf16":
            return True
        if self.conditioner_dtype != "auto":
            return False
        return tc_eligible(*shape) or (n_chains is not None and n_chains >= self.NEUTRA_AUTO_MIN_CHAINS)

    def descriptor(self, device: torch.device):
        blob = self.blob(device)
        M, H = self.conditioner_shape()
        return N.RealNVPDesc(self.n_dim, self.n_coupling, M, H, blob.data_ptr(), blob.numel()), blob

    # -- operators ----------------------------------------------------------------------------------------------
    def _pass(self, fn_name: str, x: torch.Tensor):
        dev = N.require_cuda(x.device if x.is_cuda else None)
        batch = x.shape[: x.ndim - len(self.event_shape)]
        xd = N.dev_f32(x, dev).reshape(-1, self.n_dim)
        n = xd.shape[0]
        y = torch.empty_like(xd)
        ld = torch.empty(n, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):           # the C entry points launch on the current device
            if self.uses_tensor_cores():
                desc, keep = self.tc_descriptor(dev)
                mode = 0 if fn_name == "nfmc_realnvp_forward" else 1
                N.check(N.lib().nfmc_flow_tc_pass(C.byref(desc), mode, N.ptr(xd), N.ptr(y), N.ptr(ld), n, N.stream_ptr(dev)))
            elif self.uses_row_tile_pass():
                desc, keep = self.theta_descriptor(dev)
                N.check(N.lib().nfmc_flow_wide_pass(desc.d, desc.n_coupling, desc.n_linear, desc.hidden, N.ptr(keep),
                                                    2 if fn_name == "nfmc_realnvp_forward" else 3, N.ptr(xd), N.ptr(y), N.ptr(ld), n,
                                                    N.stream_ptr(dev)))
            else:
                desc, keep = self.descriptor(dev)
                N.check(getattr(N.lib(), fn_name)(C.byref(desc), N.ptr(xd), N.ptr(y), N.ptr(ld), n, N.stream_ptr(dev)))
        return y.reshape(*batch, *self.event_shape), ld.reshape(batch)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, context=None):
        return self._pass("nfmc_realnvp_forward", x)

    @torch.no_grad()
    def inverse(self, z: torch.Tensor, context=None):
        return self._pass("nfmc_realnvp_inverse", z)


class Flow(nn.Module):
    """Standard-normal base distribution + bijection."""

    def __init__(self, bijection: RealNVP):
        super().__init__()
        if not isinstance(bijection, RealNVP):
            raise NotImplementedError("the accelerated path supports RealNVP flows only")
        self.bijection = bijection
        self.register_buffer("_device_probe", torch.zeros(()))
        self._sample_calls = 0
        self._seed = None

    @property
    def event_shape(self):
        return self.bijection.event_shape

    def get_device(self) -> torch.device:
        return self._device_probe.device

    def _compute_device(self) -> torch.device:
        d = self.get_device()
        return N.require_cuda(d if d.type == "cuda" else None)

    @torch.no_grad()
    def log_prob(self, x: torch.Tensor, context=None) -> torch.Tensor:
        dev = self._compute_device()
        bij = self.bijection
        batch = x.shape[: x.ndim - len(bij.event_shape)]
        xd = N.dev_f32(x, dev).reshape(-1, bij.n_dim)
        n = xd.shape[0]
        out = torch.empty(n, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):           # the C entry points launch on the current device
            if bij.uses_tensor_cores():
                desc, keep = bij.tc_descriptor(dev)
                N.check(N.lib().nfmc_flow_tc_pass(C.byref(desc), 2, N.ptr(xd), None, N.ptr(out), n, N.stream_ptr(dev)))
            elif bij.uses_row_tile_pass():
                desc, keep = bij.theta_descriptor(dev)
                N.check(N.lib().nfmc_flow_wide_log_prob(desc.d, desc.n_coupling, desc.n_linear, desc.hidden, N.ptr(keep), 1,
                                                        N.ptr(xd), N.ptr(out), n, N.stream_ptr(dev)))
            else:
                desc, keep = bij.descriptor(dev)
                N.check(N.lib().nfmc_flow_log_prob(C.byref(desc), N.ptr(xd), N.ptr(out), n, N.stream_ptr(dev)))
        return out.reshape(batch)

    @torch.no_grad()
    def sample(self, sample_shape, context=None, no_grad: bool = False, return_log_prob: bool = False,
               seed: Optional[int] = None, z: Optional[torch.Tensor] = None):
        """Draw from the flow.  ``seed`` keys the Philox stream (default: drawn from torch's global generator);
        ``z`` injects the base draw instead."""
        dev = self._compute_device()
        bij = self.bijection
        if isinstance(sample_shape, int):
            sample_shape = (sample_shape,)
        n = int(math.prod(sample_shape))
        x = torch.empty(n, bij.n_dim, device=dev, dtype=torch.float32)
        lq = torch.empty(n, device=dev, dtype=torch.float32)
        zd = None if z is None else N.dev_f32(z, dev).reshape(n, bij.n_dim)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (), dtype=torch.int64))
        rng = N.rng_desc(seed, 0, zd, None)
        with torch.cuda.device(dev):
            if bij.row_tile_supported():       # also tensor-core flows: there is no tcgen05 sampling entry, and the row-tile fp32
                desc, keep = bij.theta_descriptor(dev)      # kernel is several times the per-chain generic conditioner
                N.check(N.lib().nfmc_flow_wide_sample(desc.d, desc.n_coupling, desc.n_linear, desc.hidden, N.ptr(keep), 1, C.byref(rng), 0,
                                                      N.ptr(x), N.ptr(lq), n, N.stream_ptr(dev)))
            else:
                desc, keep = bij.descriptor(dev)
                N.check(N.lib().nfmc_flow_sample(C.byref(desc), C.byref(rng), 0, N.ptr(x), N.ptr(lq), n, N.stream_ptr(dev)))
        x = x.reshape(*sample_shape, *bij.event_shape)
        if return_log_prob:
            return x, lq.reshape(tuple(sample_shape))
        return x

    # -- training (nfmc_b200/flow_train.py: native kernels for the default conditioners, torch autograd otherwise) ---------
    def fit(self, x_train, *args, **kwargs):
        """Maximum-likelihood fit (reference call sites: jump.py:139-151,201; imh.py:171-175).  Raises ``ValueError`` when
        the loss becomes non-finite, which the callers turn into a weight rollback."""
        from . import flow_train
        return flow_train.fit(self, x_train, *args, **kwargs)

    def variational_fit(self, target_log_prob, *args, **kwargs):
        """Reverse-KL fit to ``target_log_prob`` (reference call sites: imh.py:67-72, neutra.py:84-91)."""
        from . import flow_train
        return flow_train.variational_fit(self, target_log_prob, *args, **kwargs)


# ------------------------------------------------------------------------------------------------------------------
# packing
# ------------------------------------------------------------------------------------------------------------------
SMALL_H = 8   # conditioners with 2 linear layers and <= 8 hidden units use the register-resident kernel path


def blob_floats(d: int, Lc: int, M: int, H: int) -> int:
    da, db = d // 2, d - d // 2
    if M == 2 and H <= SMALL_H:
        per = da * SMALL_H + SMALL_H + db * 2 * SMALL_H + ((2 * db + 3) // 4) * 4
    elif M == 1:
        per = da * 2 * db + 2 * db
    else:
        per = (da * H + H) + (M - 2) * (H * H + H) + (H * 2 * db + 2 * db)
    return (Lc + 1) * 4 * d + 4 + Lc * per


@torch.no_grad()
def pack_realnvp(bij: RealNVP) -> torch.Tensor:
    """Pack a RealNVP into the flat fp32 blob the kernels read (layout: nfmc_b200/csrc/flow.cuh header).

    Reverse permutations are folded in: after r reversals logical position p sits at physical coordinate
    ``p`` (r even) or ``d-1-p`` (r odd); every per-dimension parameter and every conditioner weight is stored
    at the physical coordinate.  Coupling l (0-based) follows r = l+1 reversals.  Elementwise affines that are
    not separated by a coupling (the last act-norm, the trailing ElementwiseAffine and its ActNorm) are composed
    into one map.
    """
    d = bij.n_dim
    da, db = d // 2, d - d // 2
    Lc = bij.n_coupling
    M, H = bij.conditioner_shape()
    small = (M == 2 and H <= SMALL_H)
    layers = list(bij.layers)
    parts = [pack_realnvp_affines(bij)]
    for l in range(Lc):
        cpl = layers[2 + 3 * l]
        if (cpl.n_linear, cpl.n_hidden) != (M, H) and cpl.n_linear > 1:
            raise ValueError("all couplings must share the conditioner shape")
        odd = (l + 1) % 2 == 1
        lin = [(m.weight.detach().to("cpu", torch.float32), m.bias.detach().to("cpu", torch.float32)) for m in cpl.linears()]
        parts += _pack_coupling_fp32(lin, da, db, M, H, small, odd)
    blob = torch.cat([p.reshape(-1) for p in parts]).contiguous()
    assert blob.numel() == blob_floats(d, Lc, M, H), (blob.numel(), blob_floats(d, Lc, M, H))
    return blob


@torch.no_grad()
def pack_realnvp_affines(bij: "RealNVP") -> torch.Tensor:
    """The elementwise-affine section of the blob: (Lc+1) tables {alpha, beta}[d], {1/alpha, -beta/alpha}[d] in
    physical coordinates, then 4 floats whose first is the total log-det constant."""
    d = bij.n_dim
    Lc = bij.n_coupling
    layers = list(bij.layers)

    def phys(layer, r):
        v = layer.value.detach().to("cpu", torch.float32)
        alpha, log_alpha, beta = _affine(v[:, 0], v[:, 1])
        if r % 2 == 1:
            alpha, beta = alpha.flip(0), beta.flip(0)
        return alpha, beta, log_alpha.sum()

    # groups of affines in forward order: group g sits before coupling g (g < Lc) / after the last coupling (g = Lc)
    groups = [[phys(layers[0], 0)]]
    for l in range(Lc):
        groups.append([phys(layers[3 + 3 * l], l + 1)])
    groups[-1] += [phys(layers[-2], Lc), phys(layers[-1], Lc)]
    parts = []
    log_const = torch.zeros((), dtype=torch.float32)
    for grp in groups:
        alpha, beta = torch.ones(d), torch.zeros(d)
        for a, b, ls in grp:                       # y = a * (alpha x + beta) + b
            alpha, beta = a * alpha, a * beta + b
            log_const = log_const + ls
        ralpha = 1.0 / alpha
        parts += [torch.stack([alpha, beta], dim=1).reshape(-1), torch.stack([ralpha, -beta * ralpha], dim=1).reshape(-1)]
    parts.append(torch.stack([log_const, torch.zeros(()), torch.zeros(()), torch.zeros(())]))
    return torch.cat([p.reshape(-1) for p in parts]).contiguous()


def _pack_coupling_fp32(lin, da, db, M, H, small, odd):
    """fp32 weights of one coupling in the kernels' layout (``lin`` = [(weight, bias), ...] of the conditioner's linear
    layers; ``odd`` = an odd number of reversals precede the coupling, so source / target indices are flipped)."""
    if small:                                            # M == 2, H <= 8: hidden-minor, zero padded to 8
        (w1, b1), (wl, bl) = lin                         # [H, da], [2*db, H]
        w1t = torch.zeros(da, SMALL_H)
        w1t[:, :H] = (w1.flip(1) if odd else w1).t()     # [ks][h]
        b1p = torch.zeros(SMALL_H)
        b1p[:H] = b1
        wlt = torch.zeros(db, 2, SMALL_H)
        wlt[:, :, :H] = wl.reshape(db, 2, H)             # [t_log][c][h]
        blt = bl.reshape(db, 2)                          # [t_log][c]
        if odd:
            wlt, blt = wlt.flip(0), blt.flip(0)
        blp = torch.zeros(((2 * db + 3) // 4) * 4)
        blp[: 2 * db] = blt.reshape(-1)
        return [w1t.reshape(-1), b1p, wlt.reshape(-1), blp]
    if M >= 2:
        w1, b1 = lin[0]                                  # [H, da]
        parts = [(w1.flip(1) if odd else w1).reshape(-1), b1]
        for wm, bm in lin[1:-1]:                         # [H_out, H_in] -> [H_in][H_out]
            parts += [wm.t().contiguous().reshape(-1), bm]
        wl, bl = lin[-1]                                 # [2*db, H]
        wl = wl.reshape(db, 2, H).permute(2, 1, 0)       # [H][2][db]
        bl = bl.reshape(db, 2).t()                       # [2][db]
        if odd:
            wl, bl = wl.flip(2), bl.flip(1)
        return parts + [wl.contiguous().reshape(-1), bl.contiguous().reshape(-1)]
    wl, bl = lin[0]                                      # M == 1: [2*db, da]
    wl = wl.reshape(db, 2, da).permute(2, 1, 0)          # [da][2][db]
    bl = bl.reshape(db, 2).t()
    if odd:
        wl, bl = wl.flip(0).flip(2), bl.flip(1)
    return [wl.contiguous().reshape(-1), bl.contiguous().reshape(-1)]


def tc_supported(d: int, M: int, H: int) -> bool:
    """Shapes the tcgen05 conditioner kernel (csrc/cond_tc.cu) can run: the hidden width is zero-padded to a multiple
    of 16 at pack time, so any H <= 256 works."""
    return M == 2 and d % 2 == 0 and 2 <= d <= 128 and 1 <= H <= 256


def tc_eligible(d: int, M: int, H: int) -> bool:
    """Shapes for which ``conditioner_dtype='auto'`` picks the tensor-core path: every 2-layer conditioner wider than the
    register-resident path covers (H > 8; the hidden width is zero-padded to a multiple of 16).  The alternative for
    9 <= H <= 15 would be the generic CUDA-core path, which is an order of magnitude slower."""
    return tc_supported(d, M, H) and H > SMALL_H


def _tc_hidden(H: int) -> int:
    """Hidden width of the packed tensor-core images: zero-padded to a multiple of 32 (the flow / jump kernels take any
    multiple of 16, the NeuTra kernel's K-step pairing needs 32 -- padding to 32 makes every tensor-core flow a NeuTra
    tensor-core flow too, e.g. a default H = 5 conditioner opted in with ``conditioner_dtype='bf16'``)."""
    return ((H + 31) // 32) * 32


@torch.no_grad()
def pack_realnvp_tc(bij: "RealNVP") -> torch.Tensor:
    """Pack a RealNVP for the tensor-core path (csrc/tc_common.cuh): fp32 affine tables (same as ``pack_realnvp``),
    then per coupling, bf16 weights in the UMMA shared-memory image ``[K/8][rows][8]`` (K-major, no swizzle):

    * ``W1 image [K1/8][Hp][8]``, K1 = d/2 + 2 rounded up to 16: columns k < d/2 hold W1, column d/2 holds b1 rounded to
      bf16 and column d/2 + 1 the rounding remainder (the kernel feeds constant ones there, so the bias rides on the GEMM
      with ~16 bits);
    * ``Wl' image [Hp/8][N2p][8]`` (row n = 2 t + c, N2p = 2*db rounded up to 16) with the scale rows multiplied by
      log2(e)/2 and the shift rows by 1/2;
    * ``bl' [N2p]`` fp32: ``(bl_a / 2 + log(1 - m)) * log2(e)`` and ``bl_b / 2``, so that the kernel's
      ``alpha = 2^(u_a') + m``, ``beta = u_b'`` equal ``exp(log(1-m) + u_a/2) + m``, ``u_b/2`` of the specification.
    Reverse permutations are folded in as in ``pack_realnvp``."""
    import math
    d = bij.n_dim
    da, db = d // 2, d - d // 2
    Lc = bij.n_coupling
    M, H = bij.conditioner_shape()
    if not tc_supported(d, M, H):
        raise ValueError("shape not supported by the tensor-core path")
    Hp = _tc_hidden(H)
    n2p = ((2 * db + 15) // 16) * 16
    k1 = ((da + 2 + 15) // 16) * 16
    log2e = 1.0 / math.log(2.0)
    c = math.log1p(-MIN_SCALE)
    base = pack_realnvp_affines(bij)
    chunks = [base.contiguous().view(torch.uint8)]
    for l in range(Lc):
        w1p, wlp, blp = _tc_coupling_matrices(bij, l)
        img1 = w1p.reshape(Hp, k1 // 8, 8).permute(1, 0, 2).contiguous()   # [kg][h][8]
        img2 = wlp.reshape(n2p, Hp // 8, 8).permute(1, 0, 2).contiguous()  # [kg][n][8]
        chunks += [img1.to(torch.bfloat16).view(torch.uint8).reshape(-1), img2.to(torch.bfloat16).view(torch.uint8).reshape(-1),
                   blp.contiguous().view(torch.uint8).reshape(-1)]
    return torch.cat([c_.reshape(-1) for c_ in chunks]).contiguous()


@torch.no_grad()
def _tc_coupling_matrices(bij: "RealNVP", l: int):
    """The padded fp32 matrices behind coupling l's tensor-core images: ``W1aug [Hp][K1]`` (W1 | b1_hi | b1_lo, zero
    padded), ``Wl' [N2p][Hp]`` (row n = 2 t + c, constants folded) and ``bl' [N2p/2][2]``."""
    import math
    d = bij.n_dim
    da, db = d // 2, d - d // 2
    M, H = bij.conditioner_shape()
    Hp = _tc_hidden(H)
    n2p = ((2 * db + 15) // 16) * 16
    k1 = ((da + 2 + 15) // 16) * 16
    log2e = 1.0 / math.log(2.0)
    c = math.log1p(-MIN_SCALE)
    cpl = list(bij.layers)[2 + 3 * l]
    odd = (l + 1) % 2 == 1
    (w1, b1), (wl, bl) = [(m.weight.detach().to("cpu", torch.float32), m.bias.detach().to("cpu", torch.float32))
                          for m in cpl.linears()]
    w1p = torch.zeros(Hp, k1)
    w1p[:H, :da] = w1.flip(1) if odd else w1                      # [h][ks], zero rows for the padded hidden units
    b1_hi = b1.to(torch.bfloat16).to(torch.float32)
    w1p[:H, da] = b1_hi
    w1p[:H, da + 1] = b1 - b1_hi
    wl3 = wl.reshape(db, 2, H).clone()                            # [t_log][c][h]
    bl2 = bl.reshape(db, 2).clone()
    if odd:
        wl3, bl2 = wl3.flip(0), bl2.flip(0)
    wl3[:, 0, :] *= 0.5 * log2e
    wl3[:, 1, :] *= 0.5
    wlp = torch.zeros(n2p, Hp)
    wlp[: 2 * db, :H] = wl3.reshape(2 * db, H)                    # row n = 2 t + c
    blp = torch.zeros(n2p // 2, 2)
    blp[:, 0] = c * log2e                                         # padded targets: alpha = 1, beta = 0
    blp[:db, 0] = (0.5 * bl2[:, 0] + c) * log2e
    blp[:db, 1] = 0.5 * bl2[:, 1]
    return w1p, wlp, blp


def neutra_tc_supported(d: int, M: int, H: int, Lc: int) -> bool:
    """Shapes the tensor-core NeuTra kernel (csrc/tc_neutra.cu) runs: tensor-core eligible, hidden width (padded to 16) a
    multiple of 32, d % 4 == 0 and a shared-memory plan that fits (the library decides)."""
    if not tc_supported(d, M, H) or Lc < 1:
        return False
    return N.lib().nfmc_neutra_tc_transposed_bytes(d, Lc, _tc_hidden(H)) > 0


@torch.no_grad()
def pack_realnvp_tc_transposed(bij: "RealNVP") -> torch.Tensor:
    """The dgrad operands of the tensor-core NeuTra kernel (csrc/tc_neutra.cu), per coupling: ``Wl'^T image
    [N2p/8][Hp][8]`` (B operand of dH = dU' . Wl': row = hidden unit, K = U' column) and ``W1^T image [Hp/8][K1][8]``
    (B operand of dS = dHpre . W1: row = source index, K = hidden unit), bf16, the same rounded values as the forward
    images of ``pack_realnvp_tc``."""
    d = bij.n_dim
    da, db = d // 2, d - d // 2
    M, H = bij.conditioner_shape()
    Hp = _tc_hidden(H)
    n2p = ((2 * db + 15) // 16) * 16
    k1 = ((da + 2 + 15) // 16) * 16
    chunks = []
    for l in range(bij.n_coupling):
        w1p, wlp, _ = _tc_coupling_matrices(bij, l)
        img_wlT = wlp.reshape(n2p // 8, 8, Hp).permute(0, 2, 1).contiguous()   # [kg over U' columns][h][8]
        img_w1T = w1p.reshape(Hp // 8, 8, k1).permute(0, 2, 1).contiguous()    # [kg over hidden units][source j][8]
        chunks += [img_wlT.to(torch.bfloat16).view(torch.uint8).reshape(-1), img_w1T.to(torch.bfloat16).view(torch.uint8).reshape(-1)]
    return torch.cat(chunks).contiguous()


def create_flow_object(flow_string: str, event_shape, **kwargs) -> Flow:
    """``'realnvp'`` or ``'realnvp%{json kwargs}'`` -> Flow (the reference's ``nfmc.util.create_flow_object``,
    /root/reference/nfmc/util.py:189-215,218-281, restricted to the RealNVP family)."""
    import json
    name, _, js = flow_string.partition("%")
    if js:
        kwargs = {**kwargs, **json.loads(js)}
    if name.lower() not in ("realnvp", "real_nvp", "rnvp"):
        raise NotImplementedError(f"flow {name!r}: only 'realnvp' is on the accelerated hot path")
    return Flow(RealNVP(event_shape, **kwargs))
