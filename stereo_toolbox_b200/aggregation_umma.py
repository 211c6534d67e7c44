"""Tensor-core backend of the 3-D aggregation: channels-last bf16 activations, tcgen05 implicit-GEMM
convolution (csrc/conv3d_umma.cu) with a CUDA-core companion for the layer shapes it does not take yet.

Same interface as aggregation.Fp32Backend, so the model code (gwcnet.py / psmnet.py) is unchanged:
tensors that flow between backend calls are [B,D,H,W,C] bf16 here.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from .aggregation import _NoProf, _bias_ver, _split, conv_work
from .ops import ACT, _p, _stream

# UMMA descriptor base-offset convention for row-shifted operand views, settled by
# csrc/probe/umma_probe.cu on a B200 (see profiles/umma_probe_r01.txt).
# profiles/umma_probe_r01.txt: shifted views work with base_offset = 0 (absolute-address swizzle).
BO_MODE = int(os.environ.get("STB_UMMA_BO_MODE", "0"))
ES_VARIANT = int(os.environ.get("STB_TMA_ES_VARIANT", "0"))      # box extent convention under TMA element strides
FORCE_SIMT = os.environ.get("STB_UMMA_FORCE_SIMT", "0") == "1"
KWMERGE = os.environ.get("STB_UMMA_KWMERGE", "1") == "1"          # merge the 3 kw taps along N (N = 3*Cout) for k3 s1 convs
DECONV_MERGE = os.environ.get("STB_UMMA_DECONV_MERGE", "1") == "1"  # transposed conv: 8 parity classes in one accumulator round
PAIRMERGE = os.environ.get("STB_UMMA_PAIRMERGE", "1") == "1"        # stride-2 convs: kw = 0 / 2 taps as one N = 2*Cn MMA
KDEPTH3D = os.environ.get("STB_UMMA_KDEPTH3D", "1") == "1"          # 3-D layers: K-chunks accumulated in TMEM when the weights fit
GRAYCLS = os.environ.get("STB_UMMA_GRAYCLS", "1") == "1"            # merged transposed convs on 16-channel slices: Gray order of the class blocks
KGROUP = os.environ.get("STB_UMMA_KGROUP", "1") == "1"              # ... and G < nk chunks per K-split pass when all of them do not
SIMT_STRIDE2 = os.environ.get("STB_UMMA_SIMT_STRIDE2", "0") == "1"   # keep strided convs on the CUDA-core companion
TORCH_DT = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp16x2": torch.float16}
SPLIT_FLAG = 64      # stb_conv3d_umma flags bit6: operand-split fp16 ("fp16x2")


def split_pack(w: torch.Tensor) -> torch.Tensor:
    """fp32 [..., C] (C % 16 == 0) -> operand-split fp16 [..., 2C]: every value x is stored as hi = fp16(x) and
    lo = fp16(x - hi) (x = hi + lo to 22 mantissa bits), interleaved per 16 channels: [hi 0..15 | lo 0..15 | hi 16..31 | ...]
    -- the storage format of the 'fp16x2' precision (csrc/conv3d_umma.cu, SPLIT)."""
    c = w.shape[-1]
    assert c % 16 == 0
    hi = w.to(torch.float16)
    lo = (w - hi.float()).to(torch.float16)
    lead = w.shape[:-1]
    return torch.stack((hi.reshape(*lead, c // 16, 16), lo.reshape(*lead, c // 16, 16)), dim=-2).reshape(*lead, 2 * c)


def split_weight_exponent(w: torch.Tensor) -> int:
    """Power of two s the (BN-folded) weights of a layer are multiplied by before split_pack: brings max|w| to [2^13, 2^14),
    so that the lo halves (|lo| <= 2^-12 |w|) of all but the smallest weights are NORMAL fp16 numbers (an fp16 subnormal lo
    carries an absolute error of 2^-25, i.e. only ~2^-20 relative to a typical 0.05-sized weight).  The kernel multiplies its
    fp32 accumulators by 2^-s (flags bits 16..22 of stb_conv3d_umma)."""
    m = float(w.abs().max())
    if not (m > 0.0) or m != m or m == float("inf"):
        return 0
    import math
    return max(0, min(100, 13 - math.floor(math.log2(m))))


def split_unpack(x: torch.Tensor) -> torch.Tensor:
    """Inverse of split_pack: fp16 [..., 2C] -> fp32 [..., C]."""
    c2 = x.shape[-1]
    v = x.reshape(*x.shape[:-1], c2 // 32, 2, 16).float()
    return (v[..., 0, :] + v[..., 1, :]).reshape(*x.shape[:-1], c2 // 2)


def pad_channels(c: int) -> int:
    """Channel count the channels-last tensors are padded to: one swizzle row (16/32/64 ch) or a multiple of 64."""
    for k in (16, 32, 64):
        if c <= k:
            return k
    return (c + 63) // 64 * 64


_TILE_IDX_CACHE: Dict[tuple, torch.Tensor] = {}


def _iarr(vals):
    return (ctypes.c_int * len(vals))(*vals)


class UmmaPlan:
    """Everything one Conv3d / ConvTranspose3d layer needs on the tensor-core path: bf16 weight tiles
    [tile][kc][Cpad][KC], the tap / class tables of stb_conv3d_umma_bf16, and the fp32 tap plan for the
    CUDA-core companion."""

    def __init__(self, conv, bn, cin_tensor: int, dtype=torch.bfloat16, split: bool = False):
        """``cin_tensor``: LOGICAL channel count of the input tensor (split tensors store 2 elements per channel)."""
        self.dtype = dtype
        self.split = split
        w = conv.weight.detach().float()
        tr = isinstance(conv, nn.ConvTranspose3d)
        stride, pad, k = conv.stride[0], conv.padding[0], conv.kernel_size[0]
        assert conv.kernel_size[0] == conv.kernel_size[1] == conv.kernel_size[2]
        cin, cout = (w.shape[0], w.shape[1]) if tr else (w.shape[1], w.shape[0])
        self.cin, self.cout, self.k, self.stride, self.pad, self.tr = cin, cout, k, stride, pad, tr
        self.cin_tensor = cin_tensor
        bnp = None if bn is None else (bn.weight, bn.bias, bn.running_mean, bn.running_var)
        eps = 1e-5 if bn is None else bn.eps
        opad = conv.output_padding[0] if tr else 0
        if cin_tensor != cin:       # zero-padded input channels (e.g. the 40-ch gwc volume padded to 48)
            wpad = torch.zeros((cin_tensor,) + tuple(w.shape[1:]) if tr else (w.shape[0], cin_tensor) + tuple(w.shape[2:]),
                               device=w.device)
            if tr:
                wpad[:cin] = w
            else:
                wpad[:, :cin] = w
            w = wpad
        # fp32 tap plan of the CUDA-core companion: built lazily (its ~30 small tensor ops per layer are pure overhead on the
        # tensor-core path, and the 16-bit training backend rebuilds every plan after every optimizer step)
        self._simt_args = (w, bnp, stride, pad, tr, opad, eps, conv.bias)
        self._simt = None
        self.opad = opad
        if bnp is not None:
            gamma, beta, mean, var = [t.detach().float() for t in bnp]
            scale = gamma / torch.sqrt(var + eps)
            self.shift = (beta - mean * scale).contiguous()
        else:
            scale, self.shift = None, None
        if conv.bias is not None:       # BN(y + b) = scale*y + (shift + scale*b)
            sb = conv.bias.detach().float() if scale is None else scale * conv.bias.detach().float()
            self.shift = sb.contiguous() if self.shift is None else (self.shift + sb).contiguous()
        self.umma_ok = (not FORCE_SIMT) and self._build_umma(w, bnp, eps)

    @property
    def simt(self):
        if self._simt is None:
            w, bnp, stride, pad, tr, opad, eps, bias = self._simt_args
            self._simt = ops.ConvPlan(w, bnp, stride, pad, tr, opad, eps, bias=bias)
        return self._simt

    def out_size(self, n):
        if self.tr:
            return (n - 1) * self.stride - 2 * self.pad + self.k + self.opad
        return (n + 2 * self.pad - self.k) // self.stride + 1

    def _build_umma(self, w, bnp, eps) -> bool:
        cin, cout, k, stride, pad, tr = self.cin_tensor, self.cout, self.k, self.stride, self.pad, self.tr
        if cin % 16 or (not self.split and cin not in (16, 32, 64) and cin % 64):
            return False            # (single 16-bit storage keeps its round-1 rule: other widths go to the CUDA-core companion)
        if tr and stride != 2:
            return False
        if not tr and stride not in (1, 2):
            return False
        if not tr and stride == 2 and SIMT_STRIDE2:
            return False
        in_stride = 1 if tr else stride
        cin_st = 2 * cin if self.split else cin           # storage elements per voxel
        # storage elements per K-chunk = one swizzled smem row: the largest of 64 / 32 / 16 that divides the row (48 channels,
        # IGEV's third level: 3 chunks of 16 resp. 32) -- split rows need whole (hi, lo) slice pairs, i.e. >= 32
        # ... and for which the weight tiles of the narrowest (16-channel) output slice fit next to a minimal plane ring
        # (k4 transposed convs stage 64 tiles: 131 KB at a 64-element chunk -- they take 32)
        def _fits(d):
            ring = 4 * (4 + 2) * 32 * d * 2 * (4 if in_stride == 2 else 1)
            return k ** 3 * 16 * d * 2 + ring + 4096 <= 227 * 1024
        kc = next((d for d in ((32, 16) if in_stride == 2 else (64, 32, 16)) if cin_st % d == 0 and (_fits(d) or d == 16)), 0)
        if kc == 0 or (self.split and kc < 32):
            return False
        nk = cin_st // kc
        cpad = (cout + 15) // 16 * 16

        def kdepth_group(ntaps_, ntiles_, window_, maxdh_, cblocks_, deconv_):
            """(G, full): G K-chunks per pass accumulate in TMEM along the pseudo-depth axis (0: classic K-split passes), on the
            whole padded channel count (full) or on 16-channel output slices."""
            if not (KDEPTH3D and self.split and nk > 1 and in_stride == 1):
                return 0, False
            slot = (4 + maxdh_) * 32 * kc * 2
            fits = lambda cn, g: (3072 + ((ntiles_ * g * cn * kc * 2 + 1023) & ~1023) + 1024 + (window_ * g + 1) * slot
                                  <= 227 * 1024)
            for g in range(nk, 1, -1):
                if nk % g or window_ * g > 6 or ntaps_ * g > 64:
                    continue
                ok_full = cpad * cblocks_ <= 256 and fits(cpad, g)
                ok_16 = deconv_ and cpad % 16 == 0 and cout == cpad and fits(16, g)
                # fewer passes than chunks: only where the sliced transposed-conv epilogue takes the partial (ok_16 alone)
                if (ok_full or ok_16) if g == nk else (KGROUP and ok_16 and not ok_full and g <= 7):
                    return g, ok_full
            return 0, False
        if bnp is not None:
            gamma, beta, mean, var = [t.detach().float() for t in bnp]
            scale = gamma / torch.sqrt(var + eps)
        else:
            scale = torch.ones(cout, device=w.device)
        # [kd,kh,kw,co,ci] with the BN scale folded, padded to cpad rows -> [tile][nk][cpad][KC]
        wt = (w.permute(2, 3, 4, 1, 0) if tr else w.permute(2, 3, 4, 0, 1)) * scale.view(1, 1, 1, -1, 1)
        flat = lambda a, b, c: (a * k + b) * k + c
        dz, dh, dw, sub, widx, tb, te, od0, oh0, ow0 = [], [], [], [], [], [], [], [], [], []
        nblk, cls0 = None, None
        tile_src = list(range(k ** 3))          # which (kd,kh,kw) weight matrix each staged tile holds
        self.deconv_merge = False
        self.class_order = list(range(8))       # merged transposed conv: class (cd*4 + ch*2 + cw) held by column block p
        if tr and DECONV_MERGE and stride == 2:
            # ---- merged transposed conv: one accumulator round holds all 8 output-parity classes (column block
            # c = cd*4+ch*2+cw).  A shift (a,b,c) of the input tile feeds every class whose per-dim tap set contains it;
            # each maximal run of consecutive class indices is ONE MMA (N = run*Cn) against consecutively staged tiles.
            per_dim = [[(kk, (c + pad - kk) // stride) for kk in range(k) if (c + pad - kk) % stride == 0]
                       for c in range(stride)]
            offs = sorted({o for lst in per_dim for _, o in lst})
            mn = min(offs)
            kof = [{o: kk for kk, o in per_dim[c]} for c in range(stride)]       # class bit -> {offset: k}
            shifts = [(a, b, c) for a in offs for b in offs for c in offs]
            def classes_of(sh):
                return [cd * 4 + ch * 2 + cw for cd in range(2) for ch in range(2) for cw in range(2)
                        if sh[0] in kof[cd] and sh[1] in kof[ch] and sh[2] in kof[cw]]
            shifts.sort(key=lambda sh: -len(classes_of(sh)))                      # the all-classes shift first
            if len(classes_of(shifts[0])) == 8:
                def build_merged(order):
                    # column block p holds class order[p]; a run of consecutive POSITIONS is one MMA
                    nonlocal tile_src, nblk, cls0
                    pos = {c: p_ for p_, c in enumerate(order)}
                    tile_src, nblk, cls0 = [], [], []
                    del dz[:], dh[:], dw[:], sub[:], widx[:]
                    for sh in shifts:
                        ps = sorted(pos[c] for c in classes_of(sh))
                        i = 0
                        while i < len(ps):
                            j = i
                            while j + 1 < len(ps) and ps[j + 1] == ps[j] + 1:
                                j += 1
                            dz.append(sh[0]); dh.append(sh[1] - mn); dw.append(sh[2] - mn); sub.append(0)
                            widx.append(len(tile_src)); nblk.append(j - i + 1); cls0.append(ps[i])
                            for p_ in ps[i:j + 1]:
                                c = order[p_]
                                cd, ch, cw = c >> 2, (c >> 1) & 1, c & 1
                                tile_src.append(flat(kof[cd][sh[0]], kof[ch][sh[1]], kof[cw][sh[2]]))
                            i = j + 1
                    self.class_order = list(order)
                tb.append(0)
                build_merged(list(range(8)))
                te.append(len(dz)); od0.append(0); oh0.append(0); ow0.append(0)
                self.in_off, self.out_stride, self.merge, self.deconv_merge = mn, stride, False, True
                # Gray order of the class blocks (flags bit14; w-pairs stay adjacent, every other pair reversed): the class sets of
                # the 8 input shifts fall into 10 runs instead of 14, i.e. 10 MMAs per K-step instead of 14.  Taken where the MMAs
                # are narrow (16-channel slices of the pseudo-depth form: ncu shows the tensor pipe's operand fetch 70 % busy at
                # 22 % math) and the epilogue flavours read the order (generic, LEAN 6 / 9)
                g_, full_ = kdepth_group(len(dz), len(tile_src), max(dz) - min(dz) + 1, max(dh), 8, True)
                if GRAYCLS and g_ and not full_:
                    build_merged([0, 1, 3, 2, 6, 7, 5, 4])
                    te[-1] = len(dz)
        # ---- stride-2 pair merge: the kw = 0 and kw = 2 taps read the SAME w-parity sub-tile one position apart, so they run as
        # one MMA against the two contiguous weight tiles (N = 2*Cn instead of two N = Cn MMAs: the N = 32 MMAs of these layers
        # are operand-bandwidth bound, 40 clk for 16 clk of math) and the epilogue realigns block 1 by one lane (flags bit7)
        self.pair_merge = bool(PAIRMERGE and not tr and in_stride == 2 and k == 3 and pad == 1)
        if self.pair_merge:
            e = [kk - pad for kk in range(k)]
            par = [x % in_stride for x in e]
            off = [(x - p_) // in_stride for x, p_ in zip(e, par)]
            mn = min(off)
            assert par[0] == par[2] == 1 and off[2] - off[0] == 1 and par[1] == 0
            tile_src, nblk, cls0 = [], [], []
            for a in range(k):
                for b in range(k):
                    dz.append(e[a]); dh.append(off[b] - mn); dw.append(off[0] - mn); sub.append(par[b] * 2 + 1)
                    widx.append(len(tile_src)); nblk.append(2); cls0.append(0)
                    tile_src += [flat(a, b, 0), flat(a, b, 2)]
                    dz.append(e[a]); dh.append(off[b] - mn); dw.append(off[1] - mn); sub.append(par[b] * 2 + 0)
                    widx.append(len(tile_src)); nblk.append(1); cls0.append(0)
                    tile_src.append(flat(a, b, 1))
        full = torch.zeros(k * k * k, cpad, cin, device=w.device)
        full[:, :cout] = wt.reshape(k * k * k, cout, cin)
        # (identity order: no gather; otherwise the index tensor is cached per device -- building it from a Python list is a
        #  pageable host-to-device copy, i.e. a stream synchronisation per plan, and training rebuilds every plan every step)
        if tile_src == list(range(k ** 3)):
            tiles = full
        else:
            key = (tuple(tile_src), str(w.device))
            idx = _TILE_IDX_CACHE.get(key)
            if idx is None:
                idx = _TILE_IDX_CACHE[key] = torch.tensor(tile_src, device=w.device)
            tiles = full.index_select(0, idx)
        self.wexp = 0
        if self.split:
            self.wexp = split_weight_exponent(tiles)
            tiles = split_pack(tiles * float(2.0 ** self.wexp))   # [tile][cpad][2*cin]: (hi, lo) k-slices per 16 input channels
        self.wt = tiles.view(len(tile_src), cpad, nk, kc).permute(0, 2, 1, 3).contiguous().to(self.dtype)
        self.nwtiles, self.kc, self.nk, self.in_stride = len(tile_src), kc, nk, in_stride
        if self.nwtiles * 16 * kc * 2 > 150 * 1024 or len(dz) > 64:
            return False
        if self.deconv_merge:
            pass
        elif not tr:
            # input index = stride*o + (kk - pad); for stride 2 split into parity p and half-res offset
            e = [kk - pad for kk in range(k)]
            par = [x % in_stride for x in e]
            off = [(x - p) // in_stride for x, p in zip(e, par)]
            mn = min(off)
            tb.append(0)
            self.merge = KWMERGE and in_stride == 1 and k == 3 and pad == 1
            for a in range(0 if self.pair_merge else k):
                for b in range(k):
                    if self.merge:
                        # one MMA per (kd,kh) against the 3 contiguous weight tiles (kd,kh,0..2): N = 3*Cout;
                        # the epilogue realigns the three column blocks by 0/1/2 lanes (see conv3d_umma.cu)
                        dz.append(e[a]); dh.append(off[b] - mn); dw.append(0); sub.append(0); widx.append(flat(a, b, 0))
                        continue
                    for c in range(k):
                        dz.append(e[a]); dh.append(off[b] - mn); dw.append(off[c] - mn)
                        sub.append(par[b] * 2 + par[c] if in_stride == 2 else 0); widx.append(flat(a, b, c))
            te.append(len(dz)); od0.append(0); oh0.append(0); ow0.append(0)
            self.in_off = mn
            self.out_stride = 1
        else:
            self.merge = False
            per_dim = [[(kk, (c + pad - kk) // stride) for kk in range(k) if (c + pad - kk) % stride == 0]
                       for c in range(stride)]
            mn = min(o for lst in per_dim for _, o in lst)
            for cd in range(stride):
                for ch in range(stride):
                    for cw in range(stride):
                        tb.append(len(dz))
                        for kd, a in per_dim[cd]:
                            for kh, b in per_dim[ch]:
                                for kw, c in per_dim[cw]:
                                    dz.append(a); dh.append(b - mn); dw.append(c - mn); sub.append(0)
                                    widx.append(flat(kd, kh, kw))
                        te.append(len(dz)); od0.append(cd); oh0.append(ch); ow0.append(cw)
            self.in_off = mn
            self.out_stride = stride
        if max(dh) > 3 or max(dw) > 3 or len(dz) > 64:
            return False
        # ---- all K-chunks of the layer in ONE launch (conv3d_umma flags bit5, 3-D form): pseudo-plane = depth*nk + chunk, the
        # taps of chunk c carry dz*nk + c, so the chunks accumulate in TMEM instead of chaining K-split passes through an fp32
        # partial (64->32 s2T on fp16x2: two passes, 1.47 GB of partial written and read back per call).  Needs the weight tiles
        # of ALL chunks of an output-channel slice resident next to a ring of (window*nk + 1) chunk slots; the merged
        # transposed conv may narrow the slice to 16 channels (own epilogue instantiation, LEAN 6).
        # When all nk chunks do not fit (128->64 s2T: 4 chunks = 221 KB of weight tiles per 16-channel slice), G < nk chunks per
        # pass do: nk / G K-split passes instead of nk (flags bits 11..13; the merged transposed conv's 16-channel-slice epilogue
        # reads / writes the fp32 partial, LEAN 9).
        self.kdepth, self.kgroup = False, 0
        if len(tb) == 1:
            G, _ = kdepth_group(len(dz), self.nwtiles, max(dz) - min(dz) + 1, max(dh),
                                8 if self.deconv_merge else (3 if self.merge else 1), self.deconv_merge)
            if G:
                n0 = len(dz)
                self.kgroup = G if G < nk else 0
                base = (list(dz), list(dh), list(dw), list(sub), list(widx), None if nblk is None else list(nblk),
                        None if cls0 is None else list(cls0))
                dz, dh, dw, sub, widx = [], [], [], [], []
                nblk = None if base[5] is None else []
                cls0 = None if base[6] is None else []
                for c in range(G):
                    for t in range(n0):
                        dz.append(base[0][t] * G + c); dh.append(base[1][t]); dw.append(base[2][t]); sub.append(base[3][t])
                        widx.append(c * self.nwtiles + base[4][t])
                        if nblk is not None:
                            nblk.append(base[5][t]); cls0.append(base[6][t])
                te = [len(dz)]
                # tiles ordered [chunk][tile] (= [pass][chunk in pass][tile]): runs of consecutive tiles (merged taps) stay contiguous
                self.wt = tiles.view(self.nwtiles, cpad, nk, kc).permute(2, 0, 1, 3).contiguous().to(self.dtype)
                self.nwtiles *= G               # tiles of ONE pass
                self.kdepth = True
        self.ntaps, self.nclass = len(dz), len(tb)
        self.c_dz, self.c_dh, self.c_dw, self.c_sub, self.c_widx = _iarr(dz), _iarr(dh), _iarr(dw), _iarr(sub), _iarr(widx)
        self.c_nblk = _iarr(nblk) if nblk is not None else None
        self.c_cls0 = _iarr(cls0) if cls0 is not None else None
        self.c_tb, self.c_te, self.c_od0, self.c_oh0, self.c_ow0 = _iarr(tb), _iarr(te), _iarr(od0), _iarr(oh0), _iarr(ow0)
        self.cpad = cpad
        return True


class UmmaBackend:
    """precision 'bf16' or 'fp16': 16-bit channels-last activations, fp32 accumulation in TMEM.
    precision 'fp16x2': the exact tensor-core path.  Every activation and weight is stored operand-split (fp16 hi + fp16 lo,
    ``split_pack``), a K=16 step is three MMAs (hi*hi + hi*lo + lo*hi) into the same fp32 accumulator: fp32-level accuracy
    (<= 1e-3 px vs the fp32 reference) at about a third of the single-fp16 tensor rate.  Tensors between backend calls are
    [B,D,H,W,2C] fp16 then."""

    def __init__(self, precision: str = "bf16"):
        self.name = precision
        self.dtype = TORCH_DT[precision]
        self.split = precision == "fp16x2"
        self.f16 = int(precision in ("fp16", "fp16x2"))
        self.fmt = 2 if self.split else self.f16           # storage format code of the layout / volume entry points
        self.cmul = 2 if self.split else 1                 # storage elements per logical channel
        self._plans: Dict[int, tuple] = {}
        self._ws: Dict[tuple, torch.Tensor] = {}
        self.prof = _NoProf()
        self.dchunk = 0

    def _plan(self, layer, cin_tensor) -> UmmaPlan:
        conv, bn = _split(layer)
        ver = (conv.weight.data_ptr(), conv.weight._version, cin_tensor) + _bias_ver(conv) + \
              (() if bn is None else (bn.weight._version, bn.bias._version, bn.running_mean._version,
                                      bn.running_var._version, bn.running_mean.data_ptr()))
        hit = self._plans.get(id(conv))
        if hit is not None and hit[0] == ver:
            return hit[1]
        plan = UmmaPlan(conv, bn, cin_tensor, self.dtype, self.split)
        if self.split and not plan.umma_ok:
            raise NotImplementedError("fp16x2 precision: this layer shape is not covered by the tcgen05 kernel")
        self._plans[id(conv)] = (ver, plan)
        return plan

    def _workspace(self, numel, device):
        """fp32 partial-sum buffer of the K-split passes (grown on demand, reused across layers: launches on
        one stream are ordered, and every K-split sequence fully rewrites the region it reads)."""
        key = (device.index,)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < numel:
            ws = torch.empty(numel, device=device, dtype=torch.float32)
            self._ws[key] = ws
        return ws

    # ---------------------------------------------------------------- volumes (layout entry)
    def volume_gwc_concat(self, gwc_l, gwc_r, cat_l, cat_r, maxdisp4, groups):
        B, Cg, H, W = gwc_l.shape
        cc = 0 if cat_l is None else cat_l.shape[1]
        ct_pad = pad_channels(groups + 2 * cc)
        f = lambda t: None if t is None else ops._f32c(t)
        gwc_l, gwc_r, cat_l, cat_r = f(gwc_l), f(gwc_r), f(cat_l), f(cat_r)
        vol = torch.empty(B, maxdisp4, H, W, ct_pad * self.cmul, device=gwc_l.device, dtype=self.dtype)
        with self.prof.bracket("volume_cl16", 0.0, 4.0 * 2 * (gwc_l.numel() + (0 if cat_l is None else cat_l.numel()))
                               + 2.0 * vol.numel()):
            _lib.call("stb_volume_cl16", _p(gwc_l), _p(gwc_r), _p(cat_l), _p(cat_r), _p(vol), self.fmt, B, Cg, groups,
                      cc, H, W, maxdisp4, ct_pad, 1, _stream())
        return vol

    def volume_from_cl(self, cl, maxdisp4, groups):
        """gwc (+ concat) volume straight from the tensor-core extractor's channels-last 16-bit outputs
        (features_umma.UmmaGwcFeatures, ``cl`` = {feats: [[1,2B,h,w,C_i], ...], cat: [1,2B,h,w,Cc_pad] or None, B, cc}):
        no NCHW fp32 copy of the 320-channel feature and no torch.cat of layer2/3/4."""
        feats, cat, B, cc = cl["feats"], cl["cat"], cl["B"], cl["cc"]
        _, N, H, W, _ = feats[0].shape
        assert N == 2 * B and all(f.dtype == self.dtype and f.is_contiguous() for f in feats)
        ct_pad = pad_channels(groups + 2 * cc)
        vol = torch.empty(B, maxdisp4, H, W, ct_pad * self.cmul, device=feats[0].device, dtype=self.dtype)
        ptrs = (ctypes.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
        chs = _iarr([f.shape[-1] // self.cmul for f in feats])
        nbytes = 2.0 * (sum(f.numel() for f in feats) + (0 if cat is None else cat.numel()) + vol.numel())
        with self.prof.bracket("volume_cl16", 0.0, nbytes):
            _lib.call("stb_volume_cl16_from_cl16", ptrs, chs, len(feats), _p(cat), 0 if cat is None else cat.shape[-1] // self.cmul,
                      _p(vol), self.fmt, B, groups, cc, H, W, maxdisp4, ct_pad, 1, _stream())
        return vol

    def volume_concat(self, l, r, maxdisp4, mask_left=True, att_prob=None):
        if att_prob is not None:
            raise NotImplementedError("attention-weighted concat volume is only built on the fp32 path yet")
        B, C, H, W = l.shape
        ct_pad = pad_channels(2 * C)
        l, r = ops._f32c(l), ops._f32c(r)
        vol = torch.empty(B, maxdisp4, H, W, ct_pad * self.cmul, device=l.device, dtype=self.dtype)
        with self.prof.bracket("volume_cl16", 0.0, 4.0 * 2 * l.numel() + 2.0 * vol.numel()):
            _lib.call("stb_volume_cl16", _p(None), _p(None), _p(l), _p(r), _p(vol), self.fmt, B, 0, 0, C, H, W,
                      maxdisp4, ct_pad, int(mask_left), _stream())
        return vol

    # ---------------------------------------------------------------- conv family
    def conv(self, layer, x, act="none", residual=None):
        assert x.dtype == self.dtype and x.is_contiguous() and x.dim() == 5
        B, Di, Hi, Wi, Cst = x.shape                      # Cst: storage elements per voxel (2 per channel when split)
        Cin = Cst // self.cmul
        plan = self._plan(layer, Cin)
        Do, Ho, Wo = plan.out_size(Di), plan.out_size(Hi), plan.out_size(Wi)
        out_fp32 = plan.cout < 8          # the 32->1 classifier feeds the fp32 soft-argmin head directly
        cout_t = plan.cout if (out_fp32 or not self.split) else (plan.cout + 15) // 16 * 16
        alloc = torch.zeros if cout_t != plan.cout else torch.empty      # padded split channels must read as zero
        out = alloc(B, Do, Ho, Wo, cout_t * (1 if out_fp32 else self.cmul), device=x.device,
                    dtype=torch.float32 if out_fp32 else self.dtype)
        post_res, post_act = None, None
        if residual is not None:
            assert residual.shape == out.shape and residual.is_contiguous()
            if out_fp32:
                # fp32 output (classifier logits): the kernel's residual port is 16-bit, so the fp32 residual of PSMNet's
                # cumulative heads (cost2 = classif2 + cost1, stackhourglass.py:134-136) is added afterwards on the small
                # fp32 tensor instead of being rounded to 16 bits first
                post_res, post_act, residual, act = residual, act, None, "none"
            elif residual.dtype != self.dtype:
                assert not self.split
                residual = residual.to(self.dtype)
        fam = "conv3d_umma" if plan.umma_ok else "conv3d_taps_cl16"
        fl, by = 0.0, 0.0
        if self.prof.enabled:
            fl, by = conv_work(plan.simt, (B, Cin, Di, Hi, Wi), (B, plan.cout, Do, Ho, Wo), 2 * self.cmul, residual is not None)
        detail = ""
        if self.prof.enabled:
            detail = f"{Cin}->{plan.cout} k{plan.k} s{plan.stride}{'T' if plan.tr else ''} @{Di}x{Hi}x{Wi}"
        with self.prof.bracket(fam, fl, by, detail=detail):
            if plan.umma_ok:
                nsteps, nh, nw = (Di, Hi, Wi) if plan.tr else (Do, Ho, Wo)
                ws = self._workspace(B * Do * Ho * Wo * cout_t, x.device) if (plan.nk > 1 and (not plan.kdepth or plan.kgroup)) else None
                _lib.call("stb_conv3d_umma", _p(x), _p(plan.wt), _p(plan.shift), _p(residual), _p(out), _p(ws),
                          self.f16, B, Cst, plan.kc, Di, Hi, Wi, cout_t, plan.cout, Do, Ho, Wo, plan.ntaps,
                          plan.c_dz, plan.c_dh, plan.c_dw, plan.c_sub, plan.c_widx, plan.c_nblk, plan.c_cls0,
                          plan.nwtiles, plan.nclass,
                          plan.c_tb, plan.c_te, plan.c_od0, plan.c_oh0, plan.c_ow0, plan.in_stride, plan.out_stride,
                          nsteps, nh, nw, plan.in_off, plan.in_off, ACT[act], int(out_fp32),
                          BO_MODE | (ES_VARIANT << 1) | (4 if plan.merge else 0) | (8 if plan.deconv_merge else 0)
                          | (32 if plan.kdepth else 0) | (plan.kgroup << 11) | ((1 << 14) if plan.class_order != list(range(8)) else 0) | (128 if plan.pair_merge else 0) | ((SPLIT_FLAG | (plan.wexp << 16)) if self.split else 0),
                          self.dchunk, _stream())
            else:
                assert not self.split
                sp = plan.simt
                for sel, dd, dh, dw, T, in_s, out_s, (od0, oh0, ow0) in sp.classes:
                    nd = (Do - od0 + out_s - 1) // out_s
                    nh = (Ho - oh0 + out_s - 1) // out_s
                    nw = (Wo - ow0 + out_s - 1) // out_s
                    if nd <= 0 or nh <= 0 or nw <= 0:
                        continue
                    _lib.call("stb_conv3d_taps_cl16", _p(x), _p(sel), _p(sp.shift), _p(residual), _p(out),
                              int(out_fp32), self.f16, B, Cin, Di, Hi, Wi, plan.cout, Do, Ho, Wo, T, dd, dh, dw, in_s,
                              out_s, od0, oh0, ow0, nd, nh, nw, ACT[act], _stream())
        if post_res is not None:
            out = out + post_res.float()
            if post_act == "relu":
                out = torch.relu_(out)
            elif post_act != "none":
                raise NotImplementedError(f"activation {post_act!r} after an fp32 residual")
        return out

    # ---------------------------------------------------------------- ACVNet helpers
    def from_ncdhw(self, x):
        """fp32 [B,C,D,H,W] -> channels-last 16-bit with the channel count padded to a swizzle row."""
        return to_channels_last(x, pad_channels(x.shape[1]), self.dtype, split=self.split)

    def cost_ncdhw(self, cost):
        B, D, H, W, C = cost.shape
        assert C == 1 and cost.dtype == torch.float32
        return cost.view(B, 1, D, H, W)

    def cost_native(self, cost):
        B, _, D, H, W = cost.shape
        return cost.view(B, D, H, W, 1)

    def block_attention(self, qkv, bias, heads, block):
        if self.split:
            # split storage: the windowed attention core runs on its fp32 NCDHW instantiation (the tensors at the bottom of
            # the hourglass are 1/16-resolution and small); layout kernels on both sides
            q32 = from_channels_last(qkv, split=True)
            with self.prof.bracket("block_attention", 0.0, 4.0 * (q32.numel() + q32.numel() // 3)):
                o32 = ops.block_attention(q32, bias, heads, block, channels_last=False)
            return to_channels_last(o32, pad_channels(o32.shape[1]), self.dtype, split=True)
        B, D, H, W, C3 = qkv.shape
        with self.prof.bracket("block_attention", 4.0 * B * D * H * W * (C3 // 3) * block[0] * block[1] * block[2],
                               2.0 * (qkv.numel() + qkv.numel() // 3)):
            return ops.block_attention(qkv, bias, heads, block, channels_last=True)

    # ---------------------------------------------------------------- IGEV / CFNet helpers
    def gate(self, x, gate_logits):
        return ops.feature_gate(x, gate_logits, channels_last=True, channels=gate_logits.shape[1], split=self.split)

    def cat(self, xs):
        """channel concatenation of channels-last tensors, re-padded to a legal K width."""
        xs = list(xs)
        c = sum(t.shape[-1] for t in xs) // self.cmul
        cp = pad_channels(c)
        assert not self.split or all((t.shape[-1] // 2) % 16 == 0 for t in xs)    # whole (hi, lo) blocks only
        if cp != c:
            xs.append(torch.zeros(xs[0].shape[:-1] + ((cp - c) * self.cmul,), device=xs[0].device, dtype=xs[0].dtype))
        return torch.cat(xs, dim=-1)

    def to_ncdhw(self, x, channels=None):
        if x.dtype == torch.float32:            # [B,D,H,W,C] fp32 (classifier-style outputs)
            return x.permute(0, 4, 1, 2, 3).contiguous()
        return from_channels_last(x, channels, split=self.split)

    # ---------------------------------------------------------------- head (layout exit)
    def head(self, cost, maxdisp, H, W, align_corners=False):
        assert cost.dtype == torch.float32 and cost.shape[-1] == 1
        B, D, h, w, _ = cost.shape
        with self.prof.bracket("upsample_softargmin", 0.0, 4.0 * (cost.numel() + B * H * W)):
            return ops.upsample_softargmin(cost.view(B, D, h, w), maxdisp, H, W, align_corners)


# layout helpers for tests / callers holding reference-layout tensors
def to_channels_last(x: torch.Tensor, cpad: Optional[int] = None, dtype=torch.bfloat16, split: bool = False) -> torch.Tensor:
    """fp32 [B,C,...] -> channels-last 16-bit [B,...,cpad]; split=True: operand-split fp16 [B,...,2*cpad] (cpad % 16 == 0)."""
    x = ops._f32c(x)
    B, C = x.shape[:2]
    S = x.numel() // (B * C)
    cpad = cpad or C
    if split:
        assert dtype == torch.float16 and cpad % 16 == 0
    out = torch.empty((B,) + tuple(x.shape[2:]) + (cpad * (2 if split else 1),), device=x.device, dtype=dtype)
    _lib.call("stb_ncdhw_to_cl16", _p(x), _p(out), 2 if split else int(dtype == torch.float16), B, C, S, cpad, _stream())
    return out


def from_channels_last(x: torch.Tensor, c: Optional[int] = None, split: bool = False) -> torch.Tensor:
    assert x.dtype in (torch.bfloat16, torch.float16) and x.is_contiguous()
    B, cpad = x.shape[0], x.shape[-1] // (2 if split else 1)
    c = c or cpad
    S = x.numel() // (B * x.shape[-1])
    out = torch.empty((B, c) + tuple(x.shape[1:-1]), device=x.device, dtype=torch.float32)
    _lib.call("stb_cl16_to_ncdhw", _p(x), _p(out), 2 if split else int(x.dtype == torch.float16), B, c, S, cpad, _stream())
    return out


to_channels_last_bf16 = to_channels_last
from_channels_last_bf16 = from_channels_last
