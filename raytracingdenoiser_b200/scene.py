"""Deterministic synthetic G-buffer + 1-ray-per-pixel signals for the denoisers.

An analytic one-bounce ray caster written with torch ops so that the same code produces the (small) test inputs on the
CPU and the full-size benchmark inputs on the GPU.  It is input generation, not part of the denoiser: ground plane +
spheres + sky under a moving left-handed pin-hole camera; per pixel and frame one cosine-sampled diffuse ray, one
roughness-perturbed specular ray and one sun-disk visibility ray, seeded by an integer hash of (x, y, frame, stream).
Outputs use the reference's input encodings (front-end helpers of Shaders/Include/NRD.hlsli):
  IN_VIEWZ R32F (:sky = 1e6 > denoisingRange), IN_NORMAL_ROUGHNESS R10G10B10A2 (NRD_FrontEnd_PackNormalAndRoughness :640-667),
  IN_MV RGBA16F 2.5D motion in pixels (motionVectorScale = {1/w, 1/h, 1}), IN_DIFF/SPEC_RADIANCE_HITDIST RGBA16F
  (REBLUR_FrontEnd_PackRadianceAndNormHitDist :732-743 or RELAX_FrontEnd_PackRadianceAndHitDist :789-798),
  IN_PENUMBRA R16F (SIGMA_FrontEnd_PackPenumbra :828-834).
"""
import math

import numpy as np
import torch

MASTER_SEED = 0x4E52445F  # "NRD_"
SKY_VIEWZ = 1.0e6
HIT_DIST_PARAMS = (3.0, 0.1, 20.0, -25.0)


def _hash_u32(x):
    """PCG-style integer hash on int64 tensors holding uint32 values."""
    m = 0xFFFFFFFF
    s = (x * 747796405 + 2891336453) & m
    w = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & m
    return ((w >> 22) ^ w) & m


def _rand(ix, iy, frame, stream):
    h = _hash_u32(ix + _hash_u32(iy + _hash_u32(torch.full_like(ix, (frame * 977 + stream * 131071 + MASTER_SEED) & 0xFFFFFFFF))))
    return (h >> 8).to(torch.float32) * (1.0 / 16777216.0)


def perspective_lh(fov_y_deg, aspect, near=0.1):
    """Left-handed infinite-far projection, column-major, column vectors (clip = M * v), depth = z/w in [0,1]."""
    t = 1.0 / math.tan(math.radians(fov_y_deg) * 0.5)
    m = np.zeros((4, 4), dtype=np.float32)  # m[row, col]
    m[0, 0] = t / aspect
    m[1, 1] = t
    m[2, 2] = 1.0
    m[2, 3] = -near
    m[3, 2] = 1.0
    return m


def look_at_lh(eye, yaw, pitch):
    """worldToView for a camera at `eye` looking along yaw/pitch (left-handed, +y up)."""
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    fwd = np.array([sy * cp, sp, cy * cp], dtype=np.float64)
    right = np.cross(np.array([0.0, 1.0, 0.0]), fwd)
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    m = np.eye(4, dtype=np.float64)
    m[0, :3], m[1, :3], m[2, :3] = right, up, fwd
    m[:3, 3] = -m[:3, :3] @ np.asarray(eye, dtype=np.float64)
    return m.astype(np.float32)


def colmajor(m):
    """4x4 [row, col] -> 16 floats column-major, as CommonSettings expects."""
    return [float(v) for v in np.asarray(m, dtype=np.float32).T.reshape(-1)]


class Scene(object):
    def __init__(self, width, height, device="cpu", fov_y=60.0, num_spheres=24):
        self.w, self.h = width, height
        self.device = torch.device(device)
        self.fov_y = fov_y
        rs = np.random.RandomState(1234)
        c = np.zeros((num_spheres, 3), dtype=np.float32)
        c[:, 0] = rs.uniform(-9.0, 9.0, num_spheres)
        c[:, 2] = rs.uniform(3.0, 26.0, num_spheres)
        r = rs.uniform(0.5, 1.8, num_spheres).astype(np.float32)
        c[:, 1] = r * rs.uniform(0.6, 1.0, num_spheres)
        self.centers = torch.tensor(c, device=self.device)
        self.radii = torch.tensor(r, device=self.device)
        rough = rs.uniform(0.0, 1.0, num_spheres).astype(np.float32)
        rough[0], rough[1] = 0.0, 1.0
        self.sphere_rough = torch.tensor(rough, device=self.device)
        self.sphere_mat = torch.tensor(rs.randint(0, 4, num_spheres).astype(np.float32), device=self.device)
        self.sphere_color = torch.tensor(rs.uniform(0.1, 1.0, (num_spheres, 3)).astype(np.float32), device=self.device)
        self.sun_dir = torch.tensor(np.array([0.35, 0.8, -0.45], dtype=np.float32) / np.linalg.norm([0.35, 0.8, -0.45]), device=self.device)
        self.proj = perspective_lh(fov_y, width / float(height))
        ys, xs = torch.meshgrid(torch.arange(height, device=self.device), torch.arange(width, device=self.device), indexing="ij")
        self.ix, self.iy = xs.to(torch.int64), ys.to(torch.int64)

    # ---- camera path: slow dolly + yaw, a few pixels of parallax per frame at 1080p ----
    def camera(self, frame):
        eye = [0.6 * math.sin(frame * 0.05), 1.7 + 0.05 * math.sin(frame * 0.11), -4.0 + 0.035 * frame]
        yaw = 0.10 * math.sin(frame * 0.035)
        pitch = -0.36  # horizon at ~17 % of the frame height: most of the frame is geometry the denoiser has to process
        return look_at_lh(eye, yaw, pitch), np.array(eye, dtype=np.float32)

    # ---- analytic intersection of rays (origin o[...,3], direction d[...,3]) with the scene ----
    def _intersect(self, o, d, tmax=1.0e30):
        t_best = torch.full(o.shape[:-1], tmax, device=self.device)
        idx = torch.full(o.shape[:-1], -2, device=self.device, dtype=torch.int64)  # -2 sky, -1 plane, >=0 sphere
        # ground plane y = 0
        denom = d[..., 1]
        tp = torch.where(denom.abs() > 1e-8, -o[..., 1] / denom, torch.full_like(denom, -1.0))
        hit = (tp > 1e-4) & (tp < t_best)
        t_best = torch.where(hit, tp, t_best)
        idx = torch.where(hit, torch.full_like(idx, -1), idx)
        dd = (d * d).sum(-1)
        for i in range(self.centers.shape[0]):
            oc = o - self.centers[i]
            b = (oc * d).sum(-1)
            cc = (oc * oc).sum(-1) - self.radii[i] * self.radii[i]
            disc = b * b - dd * cc
            sq = torch.sqrt(torch.clamp(disc, min=0.0))
            t0 = (-b - sq) / dd
            t1 = (-b + sq) / dd
            t = torch.where(t0 > 1e-4, t0, t1)
            hit = (disc > 0.0) & (t > 1e-4) & (t < t_best)
            t_best = torch.where(hit, t, t_best)
            idx = torch.where(hit, torch.full_like(idx, i), idx)
        return t_best, idx

    def _surface(self, p, idx):
        """normal, roughness, material, colour at hit points."""
        n = torch.zeros_like(p)
        n[..., 1] = 1.0
        checker = ((torch.floor(p[..., 0] * 0.5) + torch.floor(p[..., 2] * 0.5)) % 2.0)
        rough = 0.15 + 0.5 * checker
        mat = torch.zeros_like(rough)
        col = torch.stack([0.35 + 0.3 * checker, 0.35 + 0.25 * checker, 0.3 + 0.2 * checker], -1)
        for i in range(self.centers.shape[0]):
            m = idx == i
            if bool(m.any()):
                ni = (p - self.centers[i]) / self.radii[i]
                n = torch.where(m[..., None], ni, n)
                rough = torch.where(m, self.sphere_rough[i].expand_as(rough), rough)
                mat = torch.where(m, self.sphere_mat[i].expand_as(mat), mat)
                col = torch.where(m[..., None], self.sphere_color[i].expand_as(col), col)
        return n, rough, mat, col

    def _sky(self, d):
        dn = d / torch.sqrt((d * d).sum(-1, keepdim=True))
        t = torch.clamp(dn[..., 1] * 0.5 + 0.5, 0.0, 1.0)
        return torch.stack([0.4 + 0.3 * t, 0.55 + 0.3 * t, 0.7 + 0.3 * t], -1) * 1.5

    def _radiance(self, p, n, idx, col):
        """very simple outgoing radiance of a hit: sun * visibility-free lambert + ambient, or sky."""
        ndl = torch.clamp((n * self.sun_dir).sum(-1), min=0.0)
        return col * (0.15 + 2.5 * ndl)[..., None]

    @staticmethod
    def _basis(n):
        s = torch.where(n[..., 2] < 0.0, -torch.ones_like(n[..., 2]), torch.ones_like(n[..., 2]))
        a = -1.0 / (s + n[..., 2])
        b = n[..., 0] * n[..., 1] * a
        t = torch.stack([1.0 + s * n[..., 0] * n[..., 0] * a, s * b, -s * n[..., 0]], -1)
        bt = torch.stack([b, s + n[..., 1] * n[..., 1] * a, -n[..., 1]], -1)
        return t, bt

    def frame(self, frame_index, radiance_mode="reblur"):
        """Returns a dict of torch tensors (H, W[, C]) in the formats listed in the module docstring plus the camera."""
        w2v, eye = self.camera(frame_index)
        w2v_prev, _ = self.camera(frame_index - 1) if frame_index > 0 else (w2v, eye)
        dev = self.device
        W, H = self.w, self.h
        tan_y = math.tan(math.radians(self.fov_y) * 0.5)
        tan_x = tan_y * W / float(H)
        u = (self.ix.to(torch.float32) + 0.5) / W
        v = (self.iy.to(torch.float32) + 0.5) / H
        dv = torch.stack([(u * 2.0 - 1.0) * tan_x, (1.0 - v * 2.0) * tan_y, torch.ones_like(u)], -1)  # view-space ray, z = 1
        rot = torch.tensor(w2v[:3, :3], device=dev)  # rows = right, up, fwd
        dw = dv @ rot  # view -> world (rot^T applied to column vector == row vector times rot)
        o = torch.tensor(eye, device=dev).expand_as(dw)
        t, idx = self._intersect(o, dw)
        sky = idx == -2
        viewz = torch.where(sky, torch.full_like(t, SKY_VIEWZ), t)
        p = o + dw * t[..., None]
        n, rough, mat, col = self._surface(p, idx)
        n = torch.where(sky[..., None], torch.tensor([0.0, 0.0, -1.0], device=dev).expand_as(n), n)

        # motion vectors (2.5D, pixels): previous uv of the same world point
        wp = torch.tensor(w2v_prev, device=dev)
        pv = p @ wp[:3, :3].T + wp[:3, 3]
        proj = torch.tensor(self.proj, device=dev)
        cx = proj[0, 0] * pv[..., 0] + proj[0, 2] * pv[..., 2]
        cy = proj[1, 1] * pv[..., 1] + proj[1, 2] * pv[..., 2]
        cw = pv[..., 2]
        u_prev = cx / cw * 0.5 + 0.5
        v_prev = cy / cw * -0.5 + 0.5
        mv = torch.stack([(u_prev - u) * W, (v_prev - v) * H, pv[..., 2] - viewz, torch.zeros_like(u)], -1)
        mv = torch.where(sky[..., None], torch.zeros_like(mv), mv)

        # ---- diffuse: cosine-weighted ray ----
        r0, r1 = _rand(self.ix, self.iy, frame_index, 1), _rand(self.ix, self.iy, frame_index, 2)
        tb, bb = self._basis(n)
        phi = r0 * (2.0 * math.pi)
        sr = torch.sqrt(r1)
        dloc = torch.stack([sr * torch.cos(phi), sr * torch.sin(phi), torch.sqrt(torch.clamp(1.0 - r1, min=0.0))], -1)
        ddir = tb * dloc[..., 0:1] + bb * dloc[..., 1:2] + n * dloc[..., 2:3]
        po = p + n * 1e-3
        td, idd = self._intersect(po, ddir, tmax=1.0e4)
        pd = po + ddir * td[..., None]
        nd, _, _, cold = self._surface(pd, idd)
        ld = torch.where((idd == -2)[..., None], self._sky(ddir), self._radiance(pd, nd, idd, cold))
        fire = _rand(self.ix, self.iy, frame_index, 3) < 0.001
        ld = torch.where(fire[..., None], ld * 50.0, ld)
        diff_hit = torch.where(idd == -2, torch.full_like(td, 1.0e4), td)

        # ---- specular: mirror direction perturbed inside a roughness-sized cone ----
        vdir = -dw / torch.sqrt((dw * dw).sum(-1, keepdim=True))
        refl = 2.0 * (vdir * n).sum(-1, keepdim=True) * n - vdir
        r2, r3 = _rand(self.ix, self.iy, frame_index, 4), _rand(self.ix, self.iy, frame_index, 5)
        ts, bs = self._basis(refl)
        ang = rough * rough * torch.sqrt(r2 / torch.clamp(1.0 - r2 * 0.98, min=1e-3)) * 0.7
        phi2 = r3 * (2.0 * math.pi)
        sdir = refl + (ts * torch.cos(phi2)[..., None] + bs * torch.sin(phi2)[..., None]) * ang[..., None]
        sdir = sdir / torch.sqrt((sdir * sdir).sum(-1, keepdim=True))
        below = (sdir * n).sum(-1) <= 0.0
        tsx, ids = self._intersect(po, sdir, tmax=1.0e4)
        ps = po + sdir * tsx[..., None]
        ns, _, _, cols = self._surface(ps, ids)
        ls = torch.where((ids == -2)[..., None], self._sky(sdir), self._radiance(ps, ns, ids, cols))
        ls = torch.where(below[..., None], torch.zeros_like(ls), ls)
        spec_hit = torch.where(ids == -2, torch.full_like(tsx, 1.0e4), tsx)
        spec_hit = torch.where(below, torch.zeros_like(spec_hit), spec_hit)

        # ---- sun visibility for SIGMA: one ray to a jittered point of the sun disk ----
        r4, r5 = _rand(self.ix, self.iy, frame_index, 6), _rand(self.ix, self.iy, frame_index, 7)
        tan_sun = math.tan(math.radians(0.5 * 0.53))
        tsu, bsu = self._basis(self.sun_dir.expand_as(n))
        rr = torch.sqrt(r4) * tan_sun
        ldir = self.sun_dir + tsu * (rr * torch.cos(r5 * 2.0 * math.pi))[..., None] + bsu * (rr * torch.sin(r5 * 2.0 * math.pi))[..., None]
        ldir = ldir / torch.sqrt((ldir * ldir).sum(-1, keepdim=True))
        tl, idl = self._intersect(po, ldir, tmax=65504.0)
        ndl = (n * self.sun_dir).sum(-1)
        dist_occ = torch.where(idl == -2, torch.full_like(tl, 65504.0), tl)
        penumbra = torch.where(dist_occ >= 65504.0, torch.full_like(dist_occ, 65504.0), torch.clamp(dist_occ * tan_sun * 0.5, max=32768.0))
        penumbra = torch.where(ndl <= 0.0, torch.zeros_like(penumbra), penumbra)
        penumbra = torch.where(sky, torch.full_like(penumbra, 65504.0), penumbra)

        # IN_TRANSLUCENCY for SIGMA_SHADOW_TRANSLUCENCY (SIGMA_FrontEnd_PackTranslucency, NRD.hlsli:848-857): every second sphere is a
        # coloured translucent occluder, the rest are opaque; a miss transmits everything
        occ = idl.clamp(min=0).to(torch.float32)
        tcol = torch.stack([0.5 + 0.5 * torch.cos(occ * 1.3), 0.5 + 0.5 * torch.cos(occ * 2.1 + 1.0), 0.5 + 0.5 * torch.cos(occ * 2.9 + 2.0)], -1) * 0.8
        tcol = torch.where(((idl % 2) == 0)[..., None] & (idl >= 0)[..., None], tcol, torch.zeros_like(tcol))
        tcol = torch.where((dist_occ >= 65504.0)[..., None], torch.ones_like(tcol), tcol)
        tcol = torch.where((ndl <= 0.0)[..., None] & ~sky[..., None], torch.zeros_like(tcol), tcol)
        transl = torch.cat([((dist_occ >= 65504.0) | sky).to(torch.float32)[..., None], torch.clamp(tcol, 0.0, 1.0)], -1)
        transl = torch.floor(transl * 255.0 + 0.5).to(torch.uint8).contiguous()

        # optional inputs (CommonSettings::isHistoryConfidenceAvailable / isDisocclusionThresholdMixAvailable): smooth screen-space
        # patterns that move with the frame index, R8_UNORM
        fu, fv = u * 6.0 + 0.37 * frame_index, v * 4.0 - 0.21 * frame_index
        diff_conf = 0.5 + 0.5 * torch.sin(fu) * torch.cos(fv)
        spec_conf = 0.5 + 0.5 * torch.cos(fu * 1.3 + 1.0) * torch.sin(fv * 0.7)
        mix = torch.clamp(0.5 + 0.5 * torch.sin(fu * 0.5 + fv), 0.0, 1.0)
        unorm8 = lambda x: torch.floor(torch.clamp(x, 0.0, 1.0) * 255.0 + 0.5).to(torch.uint8).contiguous()

        # IN_BASECOLOR_METALNESS (isBaseColorMetalnessAvailable): the surface colour, every third object metallic
        metal = ((idx % 3) == 0).to(torch.float32) * (~sky).to(torch.float32)
        bcm = torch.cat([torch.clamp(col, 0.0, 1.0), metal[..., None]], -1)
        bcm = torch.floor(bcm * 255.0 + 0.5).to(torch.uint8).contiguous()

        out = {
            "IN_BASECOLOR_METALNESS": bcm,
            "IN_DIFF_CONFIDENCE": unorm8(diff_conf), "IN_SPEC_CONFIDENCE": unorm8(spec_conf), "IN_DISOCCLUSION_THRESHOLD_MIX": unorm8(mix),
            "IN_TRANSLUCENCY": transl,
            "IN_VIEWZ": viewz.to(torch.float32).contiguous(),
            "IN_NORMAL_ROUGHNESS": pack_normal_roughness(n, rough, mat),
            "IN_MV": mv.to(torch.float16).contiguous(),
            "IN_PENUMBRA": penumbra.to(torch.float16).contiguous(),
            "worldToView": w2v, "worldToViewPrev": w2v_prev, "viewToClip": self.proj,
        }
        if radiance_mode == "reblur":
            out["IN_DIFF_RADIANCE_HITDIST"] = pack_reblur(ld, diff_hit, viewz, torch.ones_like(rough), sky)
            out["IN_SPEC_RADIANCE_HITDIST"] = pack_reblur(ls, spec_hit, viewz, rough, sky)
        else:
            out["IN_DIFF_RADIANCE_HITDIST"] = pack_relax(ld, diff_hit, sky)
            out["IN_SPEC_RADIANCE_HITDIST"] = pack_relax(ls, spec_hit, sky)
        out["IN_SIGNAL"] = out["IN_DIFF_RADIANCE_HITDIST"]  # REFERENCE denoiser: any RGBA16F signal
        return out


def checkerboard_frame(frame, frame_index, diff_mode=2, spec_mode=2):
    """Checkerboarded copy of a frame (ReblurSettings::checkerboardMode): only the pixels with ((x ^ y ^ frameIndex) & 1) == mode keep
    their radiance, packed into the left half of the texture (column x >> 1); the right half is zero.  mode 2 = full resolution.
    Reference convention: NRDSettings.h CheckerboardMode, REBLUR_PrePass.hlsli:43-100."""
    out = dict(frame)
    for name, mode in (("IN_DIFF_RADIANCE_HITDIST", diff_mode), ("IN_SPEC_RADIANCE_HITDIST", spec_mode)):
        if mode == 2 or name not in frame:
            continue
        full = frame[name]
        h, w = full.shape[0], full.shape[1]
        assert w % 2 == 0
        y = torch.arange(h, device=full.device)
        parity = ((y ^ int(frame_index) ^ int(mode)) & 1)[:, None]            # column parity of the pixels with data in row y
        cols = 2 * torch.arange(w // 2, device=full.device)[None, :] + parity  # (h, w/2)
        packed = torch.zeros_like(full)
        packed[:, : w // 2] = torch.gather(full, 1, cols[..., None].expand(h, w // 2, full.shape[2]))
        out[name] = packed.contiguous()
    return out


def pack_normal_roughness(n, rough, mat):
    """NRD_FrontEnd_PackNormalAndRoughness for R10G10B10A2 / linear roughness -> int32 tensor holding the uint32 bits."""
    v = n / n.abs().sum(-1, keepdim=True)
    wrap_x = (1.0 - v[..., 1].abs()) * torch.where(v[..., 0] >= 0.0, torch.ones_like(v[..., 0]), -torch.ones_like(v[..., 0]))
    wrap_y = (1.0 - v[..., 0].abs()) * torch.where(v[..., 1] >= 0.0, torch.ones_like(v[..., 1]), -torch.ones_like(v[..., 1]))
    ox = torch.where(v[..., 2] >= 0.0, v[..., 0], wrap_x) * 0.5 + 0.5
    oy = torch.where(v[..., 2] >= 0.0, v[..., 1], wrap_y) * 0.5 + 0.5
    q = lambda x, m: torch.floor(torch.clamp(x, 0.0, 1.0) * m + 0.5).to(torch.int64)
    bits = q(ox, 1023.0) | (q(oy, 1023.0) << 10) | (q(rough, 1023.0) << 20) | (q(mat / 3.0, 3.0) << 30)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits)
    return bits.to(torch.int32).contiguous()


def _hit_dist_normalization(viewz, rough):
    a, b, c, d = HIT_DIST_PARAMS
    return (a + viewz.abs() * b) * (1.0 + (c - 1.0) * torch.clamp(torch.exp2(d * rough * rough), 0.0, 1.0))


def pack_reblur(radiance, hit_dist, viewz, rough, sky):
    radiance = torch.clamp(radiance, 0.0, 65504.0)
    y = radiance[..., 0] * 0.25 + radiance[..., 1] * 0.5 + radiance[..., 2] * 0.25
    co = radiance[..., 0] * 0.5 - radiance[..., 2] * 0.5
    cg = -radiance[..., 0] * 0.25 + radiance[..., 1] * 0.5 - radiance[..., 2] * 0.25
    nh = torch.clamp(hit_dist / _hit_dist_normalization(viewz, rough), 0.0, 1.0)
    o = torch.stack([y, co, cg, nh], -1)
    o = torch.where(sky[..., None], torch.zeros_like(o), o)
    return o.to(torch.float16).contiguous()


def pack_relax(radiance, hit_dist, sky):
    o = torch.cat([torch.clamp(radiance, 0.0, 65504.0), torch.clamp(hit_dist, 0.0, 65504.0)[..., None]], -1)
    o = torch.where(sky[..., None], torch.zeros_like(o), o)
    return o.to(torch.float16).contiguous()
