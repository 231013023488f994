"""Mocap clip tables: the on-disk format either side of the hot path.

Reference format (``data/mocap_data/*.txt``, read by ``MotionLib._open_all_mocap_datas``,
motion_lib.py:19-46): one JSON object per file,
``{"FrameDuration": 1/120, "LegOrder": ["FR","FL","HR","HL"], "Frames": [[x,y,z,qx,qy,qz,qw, 12 x q], ...]}``,
files taken in sorted-name order.  The engine wants all clips back to back in one
``[total_frames, 19]`` float64 table plus ``clip_offsets[n_clips+1]``.

``synthetic_mocap`` builds procedurally generated trot-like clips with the same
shape statistics as the shipped dataset (66 clips, ~229k frames at 120 Hz) for
benchmarks and GPU tests on hosts where the reference data is not available.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

MOCAP_FRAME = 19


@dataclass
class MocapTable:
    frames: np.ndarray      # [total, 19] float64
    offsets: np.ndarray     # [n_clips + 1] int32
    frame_dt: float
    names: list

    @property
    def n_clips(self):
        return len(self.offsets) - 1

    def clip(self, i):
        return self.frames[self.offsets[i]:self.offsets[i + 1]]

    def margin(self, policy_dt=0.02):
        """ML:33-35."""
        frame_rate = int(1.0 / self.frame_dt)
        return int(np.ceil(policy_dt / self.frame_dt)) + frame_rate + 2

    def validation_report(self, lower=None, upper=None):
        """Ingest check: joint-limit violations and 2*pi-like jumps per clip."""
        rep = []
        for i in range(self.n_clips):
            c = self.clip(i)
            q = c[:, 7:]
            viol = 0
            if lower is not None:
                viol = int(np.sum(np.any((q < lower) | (q > upper), axis=1)))
            jump = float(np.max(np.abs(np.diff(q, axis=0)))) / self.frame_dt if len(c) > 1 else 0.0
            qn = np.linalg.norm(c[:, 3:7], axis=1)
            rep.append({"clip": self.names[i], "frames": len(c), "limit_violations": viol,
                        "max_joint_speed": jump, "quat_norm_err": float(np.max(np.abs(qn - 1)))})
        return rep


def find_peaks_height_distance(x, height, distance):
    """numpy restatement of ``scipy.signal.find_peaks(x, height=height, distance=distance)`` as the reference uses it
    (utils/obstacle.py:16): strict local maxima (plateaus -> their middle sample), kept if >= height, then thinned so that
    no two peaks are closer than `distance` samples, highest first."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    peaks = []
    i = 1
    while i < n - 1:
        if x[i - 1] < x[i]:
            ahead = i + 1
            while ahead < n - 1 and x[ahead] == x[i]:
                ahead += 1
            if x[ahead] < x[i]:
                peaks.append((i + ahead - 1) // 2)
                i = ahead
        i += 1
    peaks = np.array([p for p in peaks if x[p] >= height], dtype=np.int64)
    if len(peaks) == 0:
        return peaks
    keep = np.ones(len(peaks), bool)
    order = np.argsort(x[peaks], kind="stable")           # scipy's _select_by_peak_distance: highest priority last
    dist = int(np.ceil(distance))
    for j in order[::-1]:
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < dist:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < len(peaks) and peaks[k] - peaks[j] < dist:
            keep[k] = False
            k += 1
    return peaks[keep]


def obstacle_table(table: "MocapTable"):
    """Per-clip hurdle placements (utils/obstacle.py:6-33, PLE:173-193): one plate per jump apex of the base height.
    Returns (rows [total, 4] = apex time, x, y, yaw ; offsets [n_clips + 1])."""
    frame_rate = int(1.0 / table.frame_dt)                                                   # ML:34
    rows, offs = [], [0]
    for i in range(table.n_clips):
        c = table.clip(i)
        try:                                   # the reference's own library call (utils/obstacle.py:16) when scipy is present
            from scipy.signal import find_peaks
            pk = find_peaks(c[:, 2], height=0.5, distance=120)[0]
        except ImportError:
            pk = find_peaks_height_distance(c[:, 2], 0.5, 120)
        for p in pk:
            x, y, z, w = c[p, 3:7] / np.linalg.norm(c[p, 3:7])
            yaw = np.arctan2(2 * (x * y + z * w), 1 - 2 * (y * y + z * z))                  # atan2(R[1,0], R[0,0]) (OBS:30-31)
            rows.append([p / frame_rate, c[p, 0], c[p, 1], yaw])
        offs.append(len(rows))
    return np.array(rows, dtype=np.float64).reshape(-1, 4), np.array(offs, dtype=np.int32)


def load_mocap(path) -> MocapTable:
    """Restates ML:19-46 (file discovery + JSON parse); ``path`` is a dir of ``*.txt`` or one file."""
    if not os.path.exists(path):
        raise FileNotFoundError("mocap data_path %r does not exist" % (path,))
    if os.path.isdir(path):
        files = [os.path.join(path, f) for f in sorted(f for f in os.listdir(path) if f.endswith("txt"))]
    else:
        files = [path]
    if not files:
        raise ValueError("no *.txt mocap clips under %r" % (path,))
    clips, names, frame_dt = [], [], None
    for f in files:
        with open(f, "r") as fh:
            d = json.load(fh)
        fr = np.asarray(d["Frames"], dtype=np.float64)
        if fr.ndim != 2 or fr.shape[1] != MOCAP_FRAME:
            raise ValueError("%s: frames must be [n, 19]" % f)
        if frame_dt is None:
            frame_dt = float(d["FrameDuration"])     # ML:33 uses the first file's duration for all
        clips.append(fr)
        names.append(os.path.basename(f))
    offsets = np.zeros(len(clips) + 1, np.int32)
    offsets[1:] = np.cumsum([len(c) for c in clips])
    return MocapTable(np.concatenate(clips, 0), offsets, frame_dt, names)


def save_packed(table: MocapTable, path):
    np.savez_compressed(path, frames=table.frames, offsets=table.offsets, frame_dt=table.frame_dt,
                        names=np.array(table.names))


def load_packed(path) -> MocapTable:
    d = np.load(path, allow_pickle=False)
    return MocapTable(d["frames"], d["offsets"].astype(np.int32), float(d["frame_dt"]), [str(s) for s in d["names"]])


# nominal standing/running pose (joint order FR FL HR HL x hip, thigh, shank)
_NOMINAL_Q = np.array([-0.028, -0.779, 1.687, -0.028, -0.778, 1.684, -0.028, -0.733, 1.567, -0.028, -0.732, 1.563])
_NOMINAL_Z = 0.334


def _quat_from_rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy], -1)


def synthetic_clip(n_frames, rng, frame_dt=1.0 / 120.0):
    """One procedurally generated trot-like clip: forward speed, yaw drift, diagonal-pair joint sinusoids."""
    t = np.arange(n_frames) * frame_dt
    speed = rng.uniform(0.0, 1.6)
    yaw_rate = rng.uniform(-0.4, 0.4)
    gait_hz = rng.uniform(1.5, 3.0)
    amp = rng.uniform(0.05, 0.35)
    yaw = yaw_rate * t + 0.1 * np.sin(2 * np.pi * 0.23 * t + rng.uniform(0, 6.28))
    vx, vy = speed * np.cos(yaw), speed * np.sin(yaw)
    x, y = np.cumsum(vx) * frame_dt, np.cumsum(vy) * frame_dt
    x -= x[0]
    y -= y[0]
    ph = 2 * np.pi * gait_hz * t
    z = _NOMINAL_Z + 0.01 * np.sin(2 * ph + rng.uniform(0, 6.28))
    roll = 0.03 * np.sin(ph + rng.uniform(0, 6.28))
    pitch = 0.03 + 0.03 * np.sin(2 * ph + rng.uniform(0, 6.28))
    quat = _quat_from_rpy(roll, pitch, yaw)
    q = np.tile(_NOMINAL_Q, (n_frames, 1))
    leg_phase = [0.0, np.pi, np.pi, 0.0]      # diagonal pairs
    for k in range(4):
        s = np.sin(ph + leg_phase[k])
        c = np.cos(ph + leg_phase[k])
        q[:, 3 * k + 0] += 0.05 * amp * s
        q[:, 3 * k + 1] += amp * s
        q[:, 3 * k + 2] += 0.8 * amp * np.maximum(c, 0.0) - 0.3 * amp * s
    return np.concatenate([x[:, None], y[:, None], z[:, None], quat, q], axis=1)


def synthetic_mocap(n_clips=66, seed=0, min_frames=900, max_frames=6060, frame_dt=1.0 / 120.0) -> MocapTable:
    """Synthetic stand-in for data/mocap_data: defaults give 66 clips, ~229k frames (17 MB as fp32)."""
    rng = np.random.default_rng(seed)
    clips = [synthetic_clip(int(rng.integers(min_frames, max_frames + 1)), rng, frame_dt) for _ in range(n_clips)]
    offsets = np.zeros(n_clips + 1, np.int32)
    offsets[1:] = np.cumsum([len(c) for c in clips])
    return MocapTable(np.concatenate(clips, 0), offsets, frame_dt, ["synthetic_%03d" % i for i in range(n_clips)])
