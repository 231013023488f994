"""Offline model compiler: ``max.urdf`` -> constant tables for the rollout engine.

The reference loads the robot through PyBullet's URDF importer
(``legged_robot.py:207-220``, flags ``URDF_MAINTAIN_LINK_ORDER |
URDF_USE_SELF_COLLISION | URDF_ENABLE_CACHED_GRAPHICS_SHAPES |
URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS``).  Bullet's source is *not* part
of the reference tree, so what the importer does to the numbers in the URDF is
restated here from its published behaviour:

* fixed joints are kept (22 child links, 12 revolute + 10 fixed);
* every link gets an *inertial frame* = ``<inertial><origin>`` composed with the
  principal-axes rotation of the URDF inertia tensor (``btMatrix3x3::diagonalize``
  cyclic Jacobi, threshold 1e-6, <= 30 sweeps) -- restated in
  :func:`bullet_diagonalize`;
* because the reference does **not** pass ``URDF_USE_INERTIA_FROM_FILE`` the
  principal moments are *replaced* by ``btCompoundShape::calculateLocalInertia``:
  the inertia of a solid box the size of the collision compound's AABB measured
  in that inertial frame (``use_urdf_inertia=False``, the default here).
  ``use_urdf_inertia=True`` keeps the URDF moments instead.
* pybullet's base pose/velocity API speaks in the base *inertial* frame, so the
  37-float robot state of the engine is expressed there.

The output is a flat float64 "model blob" whose layout is declared in
``include/llq_model_layout.h``; the CPU oracle consumes the *generic* section
(any tree), the CUDA engine consumes the *special* section (floating base + four
3-joint legs, fixed leaves folded into their parents).  Both sections are
derived here from the same parsed numbers and cross-checked in
``tests/test_model.py``.
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

# --------------------------------------------------------------------------- layout
# (mirrored by include/llq_model_layout.h; tests/test_model.py checks the two agree)
LLQ_MODEL_MAGIC = 0x4C4C5131  # "LLQ1"
HDR = 16          # header doubles
H_MAGIC, H_VERSION, H_NLINKS, H_NDOF, H_OFF_GENERIC, H_OFF_SPHERES, H_NSPHERES, H_OFF_SPECIAL, H_TOTAL, H_OFF_PROXIES, H_NPROXIES = range(11)
PROXY = 8         # proxy stride: link, x, y, z (link frame), radius, kind (0 foot, 1 wheel, 2 hip, 3 body corner, 4 handle), 0, 0
GL = 64           # generic per-link stride
G_PARENT, G_JTYPE, G_DOF = 0, 1, 2
G_JXYZ, G_JROT, G_AXIS = 3, 6, 15
G_MASS, G_COM, G_RIN, G_IDIAG = 18, 19, 22, 31
G_LOWER, G_UPPER, G_JDAMP, G_HASLIM = 34, 35, 36, 37
G_ICLINK = 38
SPH = 8           # sphere stride: link, cx, cy, cz, radius, friction, kind (0 foot, 1 wheel, 2 hip, 3 trunk corner, 5 thigh, 6 shank), 0
# special section
S_QI = 0          # 4  quaternion xyzw of R_I (body link axes <- base inertial axes)
S_BASE_M = 4      # 1
S_BASE_H = 5      # 3  first moment about the base reference point (body CoM), link axes
S_BASE_I = 8      # 6  xx xy xz yy yz zz about the reference point
S_BASE_ND = 14    # 1
S_BASE_DAMP = 15  # 3 items x 10 (mass, c[3], Ic[6])
DAMP_ITEM = 10
S_LEGS = 48       # start of per-leg data
LJ = 48           # per-joint stride
J_R, J_AXIS_IDX, J_AXIS_SIGN, J_M, J_H, J_I, J_ND, J_DAMP = 0, 3, 4, 5, 6, 9, 15, 16
J_LOWER, J_UPPER, J_HASLIM, J_JDAMP = 36, 37, 38, 39
LEG = 3 * LJ + 8  # per-leg stride: 3 joints + foot (cx,cy,cz,radius,friction,link_index,0,0)
L_FOOT = 3 * LJ
S_TOTAL = S_LEGS + 4 * LEG

LINEAR_DAMPING = 0.04   # btMultiBody default m_linearDamping  (SURVEY A.1 [M])
ANGULAR_DAMPING = 0.04  # btMultiBody default m_angularDamping
URDF_DEFAULT_MARGIN = 0.001  # gUrdfDefaultCollisionMargin


# --------------------------------------------------------------------------- math helpers
def rpy_to_matrix(rpy):
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix (R = Rz(y) Ry(p) Rx(r))."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def bullet_diagonalize(mat, threshold=1.0e-6, max_steps=30):
    """Restatement of ``btMatrix3x3::diagonalize`` (Bullet, not in the reference tree).

    Cyclic Jacobi that always annihilates the largest off-diagonal element and
    accumulates ``rot = rot * J``; returns (diag, rot) with ``mat = rot diag rot^T``.
    Eigenvalues are *not* sorted -- they stay next to the URDF axes they came from,
    which is what fixes the sign/ordering convention of the inertial frame.
    """
    eps = 2.220446049250313e-16  # SIMD_EPSILON of a BT_USE_DOUBLE_PRECISION build
    m = np.array(mat, dtype=np.float64).copy()
    rot = np.eye(3)
    step = max_steps
    while step > 0:
        p, q, r = 0, 1, 2
        mx = abs(m[0, 1])
        v = abs(m[0, 2])
        if v > mx:
            q, r, mx = 2, 1, v
        v = abs(m[1, 2])
        if v > mx:
            p, q, r, mx = 1, 2, 0, v
        t = threshold * (abs(m[0, 0]) + abs(m[1, 1]) + abs(m[2, 2]))
        if mx <= t:
            if mx <= eps * t:
                break
            step = 1
        mpq = m[p, q]
        theta = (m[q, q] - m[p, p]) / (2 * mpq)
        theta2 = theta * theta
        if theta2 * theta2 < 10.0 / eps:
            t = 1 / (theta + math.sqrt(1 + theta2)) if theta >= 0 else 1 / (theta - math.sqrt(1 + theta2))
            c = 1 / math.sqrt(1 + t * t)
            s = c * t
        else:
            t = 1 / (theta * (2 + 0.5 / theta2))
            c = 1 - 0.5 * t * t
            s = c * t
        m[p, q] = m[q, p] = 0.0
        m[p, p] -= t * mpq
        m[q, q] += t * mpq
        mrp, mrq = m[r, p], m[r, q]
        m[r, p] = m[p, r] = c * mrp - s * mrq
        m[r, q] = m[q, r] = c * mrq + s * mrp
        for i in range(3):
            mrp, mrq = rot[i, p], rot[i, q]
            rot[i, p] = c * mrp - s * mrq
            rot[i, q] = c * mrq + s * mrp
        step -= 1
    return np.array([m[0, 0], m[1, 1], m[2, 2]]), rot


def matrix_to_quat_xyzw(r):
    """Rotation matrix -> unit quaternion (x, y, z, w), w >= 0."""
    t = np.trace(r)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([(r[2, 1] - r[1, 2]) / s, (r[0, 2] - r[2, 0]) / s, (r[1, 0] - r[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(r)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + r[i, i] - r[j, j] - r[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (r[j, i] + r[i, j]) / s
        q[k] = (r[k, i] + r[i, k]) / s
        q[3] = (r[k, j] - r[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def sym6(m):
    return [m[0, 0], m[0, 1], m[0, 2], m[1, 1], m[1, 2], m[2, 2]]


# --------------------------------------------------------------------------- URDF parsing
def _floats(s, n=None):
    v = [float(x) for x in s.split()]
    if n is not None:
        assert len(v) == n, s
    return v


def parse_urdf(path):
    """Parse the subset of URDF that max.urdf uses; keeps URDF link order
    (``URDF_MAINTAIN_LINK_ORDER``: Bullet link index i == i-th non-base link)."""
    root = ET.parse(path).getroot()
    links = {}
    order = []
    for le in root.findall("link"):
        name = le.get("name")
        inertial = le.find("inertial")
        o = inertial.find("origin")
        ine = inertial.find("inertia")
        tensor = np.array([[float(ine.get("ixx")), float(ine.get("ixy")), float(ine.get("ixz"))],
                           [float(ine.get("ixy")), float(ine.get("iyy")), float(ine.get("iyz"))],
                           [float(ine.get("ixz")), float(ine.get("iyz")), float(ine.get("izz"))]])
        cols = []
        for ce in le.findall("collision"):
            co = ce.find("origin")
            g = ce.find("geometry")[0]
            c = {"xyz": _floats(co.get("xyz"), 3), "rpy": _floats(co.get("rpy"), 3), "type": g.tag}
            if g.tag == "box":
                c["size"] = _floats(g.get("size"), 3)
            elif g.tag == "cylinder":
                c["radius"], c["length"] = float(g.get("radius")), float(g.get("length"))
            elif g.tag == "sphere":
                c["radius"] = float(g.get("radius"))
            else:
                raise ValueError("unsupported collision geometry " + g.tag)
            cols.append(c)
        links[name] = {
            "name": name,
            "mass": float(inertial.find("mass").get("value")),
            "inertial_xyz": _floats(o.get("xyz"), 3),
            "inertial_rpy": _floats(o.get("rpy"), 3),
            "inertia_urdf": tensor.tolist(),
            "collisions": cols,
        }
        order.append(name)
    joints = {}
    for je in root.findall("joint"):
        child = je.find("child").get("link")
        o = je.find("origin")
        ax = je.find("axis")
        lim = je.find("limit")
        dyn = je.find("dynamics")
        joints[child] = {
            "joint_name": je.get("name"),
            "joint_type": je.get("type"),
            "parent": je.find("parent").get("link"),
            "joint_xyz": _floats(o.get("xyz"), 3),
            "joint_rpy": _floats(o.get("rpy"), 3),
            "axis": _floats(ax.get("xyz"), 3) if ax is not None else [0.0, 0.0, 0.0],
            "lower": float(lim.get("lower")) if lim is not None else 0.0,
            "upper": float(lim.get("upper")) if lim is not None else -1.0,
            "effort": float(lim.get("effort")) if lim is not None else 0.0,
            "velocity": float(lim.get("velocity")) if lim is not None else 0.0,
            "damping": float(dyn.get("damping")) if dyn is not None else 0.0,
            "friction": float(dyn.get("friction")) if dyn is not None else 0.0,
        }
    out = []
    index = {}
    for name in order:
        ln = dict(links[name])
        if name in joints:
            ln.update(joints[name])
        else:
            ln.update({"joint_name": None, "joint_type": "floating", "parent": None,
                       "joint_xyz": [0, 0, 0], "joint_rpy": [0, 0, 0], "axis": [0, 0, 0],
                       "lower": 0.0, "upper": -1.0, "effort": 0, "velocity": 0, "damping": 0, "friction": 0})
        index[name] = len(out)
        out.append(ln)
    assert out[0]["parent"] is None, "first URDF link must be the root"
    for ln in out:
        ln["parent_index"] = -1 if ln["parent"] is None else index[ln["parent"]]
        assert ln["parent_index"] < index[ln["name"]], "links must be listed parent-first"
    return out


# --------------------------------------------------------------------------- Bullet importer semantics
def _collision_aabb_in_inertial(link, r_in, com):
    """AABB (min, max) of the link's collision compound seen from its inertial frame,
    following btCompoundShape::addChildShape + child getAabb (margins:
    box exact, sphere exact, cylinder = 32-gon hull -> cached local AABB + 2x margin)."""
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for c in link["collisions"]:
        r_c = rpy_to_matrix(c["rpy"])
        t_c = np.array(c["xyz"], dtype=np.float64)
        # child transform in inertial frame: T_in^-1 * T_c
        r_rel = r_in.T @ r_c
        t_rel = r_in.T @ (t_c - com)
        if c["type"] == "box":
            he = 0.5 * np.array(c["size"])
        elif c["type"] == "cylinder":
            m2 = 2 * URDF_DEFAULT_MARGIN
            he = np.array([c["radius"] + m2, c["radius"] + m2, 0.5 * c["length"] + m2])
        elif c["type"] == "sphere":
            he = None
        if he is None:
            ext = np.full(3, c["radius"])
        else:
            ext = np.abs(r_rel) @ he
        lo = np.minimum(lo, t_rel - ext)
        hi = np.maximum(hi, t_rel + ext)
    return lo, hi


def bullet_link_properties(link, use_urdf_inertia=False):
    """(R_in, principal inertia diag) of one link as Bullet's URDF importer sets them."""
    tensor = np.array(link["inertia_urdf"], dtype=np.float64)
    if tensor[0, 1] == 0.0 and tensor[0, 2] == 0.0 and tensor[1, 2] == 0.0:
        diag, basis = np.diag(tensor).copy(), np.eye(3)
    else:
        diag, basis = bullet_diagonalize(tensor)
    px, py, pz = diag
    if px < 0 or px > py + pz or py < 0 or py > px + pz or pz < 0 or pz > px + py:
        diag, basis = np.zeros(3), np.eye(3)   # "Bad inertia tensor properties" branch
    r_in = rpy_to_matrix(link["inertial_rpy"]) @ basis
    mass = link["mass"]
    if mass and not use_urdf_inertia and link["collisions"]:
        lo, hi = _collision_aabb_in_inertial(link, r_in, np.array(link["inertial_xyz"]))
        lx, ly, lz = hi - lo
        diag = mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
    if not mass:
        diag = np.zeros(3)
    return r_in, diag


def compile_model(urdf_path, use_urdf_inertia=False, foot_friction=0.5):
    links = parse_urdf(urdf_path)
    dof = 0
    for ln in links:
        r_in, diag = bullet_link_properties(ln, use_urdf_inertia)
        ln["R_in"] = r_in.tolist()
        ln["inertia_diag"] = diag.tolist()
        ln["Ic_link"] = (r_in @ np.diag(diag) @ r_in.T).tolist()
        if ln["joint_type"] == "revolute":
            ln["dof_index"] = dof
            dof += 1
        else:
            ln["dof_index"] = -1
    model = {"source": os.path.basename(urdf_path), "use_urdf_inertia": bool(use_urdf_inertia),
             "n_dof": dof, "links": links, "foot_friction": foot_friction,
             "linear_damping": LINEAR_DAMPING, "angular_damping": ANGULAR_DAMPING}
    return model


# --------------------------------------------------------------------------- blob packing
LEG_ORDER = ["FR", "FL", "HR", "HL"]


def _link_index(model, name):
    for i, ln in enumerate(model["links"]):
        if ln["name"] == name:
            return i
    raise KeyError(name)


def _rigid_children(model, idx):
    """Indices of links rigidly attached (through fixed joints only) below link idx."""
    out = []
    for i, ln in enumerate(model["links"]):
        if ln["parent_index"] == idx and ln["joint_type"] == "fixed":
            out.append(i)
            out.extend(_rigid_children(model, i))
    return out


def _pose_in_ancestor(model, idx, anc):
    """(R, t) of link idx's frame expressed in ancestor link anc's frame, all joints at q=0
    (only used across fixed joints, where q does not enter)."""
    r, t = np.eye(3), np.zeros(3)
    while idx != anc:
        ln = model["links"][idx]
        rj, tj = rpy_to_matrix(ln["joint_rpy"]), np.array(ln["joint_xyz"], dtype=np.float64)
        r, t = rj @ r, rj @ t + tj
        idx = ln["parent_index"]
    return r, t


def _composite(model, idx, ref_point):
    """Composite spatial inertia (m, h, I_O) and damping items of moving link idx plus its
    fixed leaves, about ref_point (given in idx's link frame), in idx's link axes."""
    m_tot, h, i_o, items = 0.0, np.zeros(3), np.zeros((3, 3)), []
    for j in [idx] + _rigid_children(model, idx):
        ln = model["links"][j]
        mass = ln["mass"]
        r_j, t_j = _pose_in_ancestor(model, j, idx)
        c = r_j @ np.array(ln["inertial_xyz"]) + t_j - ref_point
        ic = r_j @ np.array(ln["Ic_link"]) @ r_j.T
        if mass == 0.0:
            continue
        m_tot += mass
        h += mass * c
        i_o += ic - mass * skew(c) @ skew(c)
        items.append([mass, *c.tolist(), *sym6(ic)])
    return m_tot, h, i_o, items


SPH_FOOT, SPH_WHEEL, SPH_HIP, SPH_CORNER, SPH_THIGH, SPH_SHANK = 0, 1, 2, 3, 5, 6


def contact_proxy_spheres(model):
    """Collision spheres that stand in for the robot's non-foot shapes (the reference loads every link with its collision geometry,
    LR:212-217): rows [articulated link, centre xyz in that link's frame, radius, 0, kind, 0] in the order both engines keep --
    4 knee wheels (cylinders r 0.028 / 0.036, as spheres), 4 hips (cylinder r 0.047), 4 thighs (one sphere on the 0.2 m box, r = half
    its larger cross-section), 8 shank spheres (two per 0.24 m box, r = half its larger cross-section), 8 trunk-box corners (r 0).
    llq_config.knee_contacts selects how many are live: 0 none, 1 the wheels, 2 all of them."""
    links = model["links"]

    def art(idx):                                     # first non-fixed ancestor and the pose of idx in it
        anc = idx
        while links[anc]["joint_type"] == "fixed":
            anc = links[anc]["parent_index"]
        return anc, _pose_in_ancestor(model, idx, anc)

    out = {k: [] for k in (SPH_WHEEL, SPH_HIP, SPH_THIGH, SPH_SHANK, SPH_CORNER)}
    for leg in LEG_ORDER:
        widx = _link_index(model, "link_%sW" % leg)
        cyl = [c for c in links[widx]["collisions"] if c["type"] == "cylinder"][0]
        anc, (r_w, t_w) = art(widx)
        out[SPH_WHEEL].append([anc, *(r_w @ np.array(cyl["xyz"]) + t_w), cyl["radius"], 0, SPH_WHEEL, 0])
        hidx = _link_index(model, "link_%s1" % leg)
        cyl = [c for c in links[hidx]["collisions"] if c["type"] == "cylinder"][0]
        out[SPH_HIP].append([hidx, *cyl["xyz"], cyl["radius"], 0, SPH_HIP, 0])
        tidx = _link_index(model, "link_%s2" % leg)
        box = [c for c in links[tidx]["collisions"] if c["type"] == "box"][0]
        size = sorted(box["size"])                    # the long side runs along the link (rpy turns the box's x onto z)
        ctr = np.array(box["xyz"], dtype=np.float64)
        out[SPH_THIGH].append([tidx, *(ctr + [0, 0, -0.05 * size[2]]), 0.5 * size[1], 0, SPH_THIGH, 0])
        sidx = _link_index(model, "link_%s3" % leg)
        box = [c for c in links[sidx]["collisions"] if c["type"] == "box"][0]
        size = sorted(box["size"])
        ctr = np.array(box["xyz"], dtype=np.float64)
        for off in (0.125 * size[2], -0.1875 * size[2]):          # z_c + 0.03, z_c - 0.045 for the 0.24 m shank box
            out[SPH_SHANK].append([sidx, *(ctr + [0, 0, off]), 0.5 * size[1], 0, SPH_SHANK, 0])
    body = links[0]
    box = [c for c in body["collisions"] if c["type"] == "box"][0]
    hx, hy, hz = 0.5 * np.array(box["size"])
    for sx in (1, -1):
        for sy in (1, -1):
            for sz in (1, -1):
                out[SPH_CORNER].append([0, box["xyz"][0] + sx * hx, box["xyz"][1] + sy * hy, box["xyz"][2] + sz * hz, 0.0, 0, SPH_CORNER, 0])
    return out[SPH_WHEEL] + out[SPH_HIP] + out[SPH_THIGH] + out[SPH_SHANK] + out[SPH_CORNER]


def pack_model(model):
    """Flat float64 blob (layout: include/llq_model_layout.h)."""
    links = model["links"]
    n = len(links)
    spheres = []
    for i, ln in enumerate(links):
        for c in ln["collisions"]:
            if c["type"] == "sphere" and ln["name"].endswith("4"):
                spheres.append([i, *c["xyz"], c["radius"], model["foot_friction"], 0, 0])
    spheres += contact_proxy_spheres(model)
    # detection-only proxy spheres used for "robot touches the PMC hurdle plate" (PLE:341-346): feet, wheels (knees), hips,
    # body-box corners -- a coarse stand-in for Bullet's exact link shapes
    proxies = []
    for i, ln in enumerate(links):
        nm = ln["name"]
        for c in ln["collisions"]:
            if nm.endswith("4") and c["type"] == "sphere":
                proxies.append([i, *c["xyz"], c["radius"], 0, 0, 0])
            elif nm.endswith("W") and c["type"] == "cylinder":
                proxies.append([i, *c["xyz"], c["radius"], 1, 0, 0])
            elif nm.endswith("1") and c["type"] == "cylinder":
                proxies.append([i, *c["xyz"], c["radius"], 2, 0, 0])
            elif "handle" in nm and c["type"] == "sphere":
                proxies.append([i, *c["xyz"], c["radius"], 4, 0, 0])
            elif nm == "body" and c["type"] == "box":
                hx, hy, hz = 0.5 * np.array(c["size"])
                for sx in (1, -1):
                    for sy in (1, -1):
                        for sz in (1, -1):
                            proxies.append([i, c["xyz"][0] + sx * hx, c["xyz"][1] + sy * hy, c["xyz"][2] + sz * hz, 0.0, 3, 0, 0])
    off_generic = HDR
    off_spheres = off_generic + n * GL
    off_special = off_spheres + len(spheres) * SPH
    off_proxies = off_special + S_TOTAL
    total = off_proxies + len(proxies) * PROXY
    blob = np.zeros(total, dtype=np.float64)
    blob[H_MAGIC], blob[H_VERSION], blob[H_NLINKS], blob[H_NDOF] = LLQ_MODEL_MAGIC, 1, n, model["n_dof"]
    blob[H_OFF_GENERIC], blob[H_OFF_SPHERES], blob[H_NSPHERES] = off_generic, off_spheres, len(spheres)
    blob[H_OFF_SPECIAL], blob[H_TOTAL] = off_special, total
    blob[H_OFF_PROXIES], blob[H_NPROXIES] = off_proxies, len(proxies)
    for k, pr in enumerate(proxies):
        blob[off_proxies + k * PROXY: off_proxies + (k + 1) * PROXY] = pr
    for i, ln in enumerate(links):
        g = blob[off_generic + i * GL: off_generic + (i + 1) * GL]
        g[G_PARENT] = ln["parent_index"]
        g[G_JTYPE] = {"floating": -1, "fixed": 0, "revolute": 1}[ln["joint_type"]]
        g[G_DOF] = ln["dof_index"]
        g[G_JXYZ:G_JXYZ + 3] = ln["joint_xyz"]
        g[G_JROT:G_JROT + 9] = rpy_to_matrix(ln["joint_rpy"]).reshape(-1)
        g[G_AXIS:G_AXIS + 3] = ln["axis"]
        g[G_MASS] = ln["mass"]
        g[G_COM:G_COM + 3] = ln["inertial_xyz"]
        g[G_RIN:G_RIN + 9] = np.array(ln["R_in"]).reshape(-1)
        g[G_IDIAG:G_IDIAG + 3] = ln["inertia_diag"]
        has_lim = ln["joint_type"] == "revolute" and ln["lower"] <= ln["upper"]
        g[G_LOWER], g[G_UPPER], g[G_JDAMP], g[G_HASLIM] = ln["lower"], ln["upper"], ln["damping"], float(has_lim)
        g[G_ICLINK:G_ICLINK + 9] = np.array(ln["Ic_link"]).reshape(-1)
    for k, s in enumerate(spheres):
        blob[off_spheres + k * SPH: off_spheres + (k + 1) * SPH] = s

    # ---- special section: floating base + 4 legs of 3 revolute joints
    sp = blob[off_special:]
    body = links[0]
    r_i = np.array(body["R_in"])
    sp[S_QI:S_QI + 4] = matrix_to_quat_xyzw(r_i)
    ref = np.array(body["inertial_xyz"], dtype=np.float64)  # base reference point = body CoM
    m, h, i_o, items = _composite(model, 0, ref)
    sp[S_BASE_M] = m
    sp[S_BASE_H:S_BASE_H + 3] = h
    sp[S_BASE_I:S_BASE_I + 6] = sym6(i_o)
    assert len(items) <= 3
    sp[S_BASE_ND] = len(items)
    for t, it in enumerate(items):
        sp[S_BASE_DAMP + t * DAMP_ITEM: S_BASE_DAMP + (t + 1) * DAMP_ITEM] = it
    for k, leg in enumerate(LEG_ORDER):
        lb = S_LEGS + k * LEG
        for j in range(3):
            idx = _link_index(model, "link_%s%d" % (leg, j + 1))
            ln = links[idx]
            assert ln["joint_type"] == "revolute" and ln["dof_index"] == 3 * k + j
            assert np.allclose(rpy_to_matrix(ln["joint_rpy"]), np.eye(3)), "actuated joints must have rpy 0"
            parent_expected = 0 if j == 0 else _link_index(model, "link_%s%d" % (leg, j))
            assert ln["parent_index"] == parent_expected
            jb = lb + j * LJ
            r = np.array(ln["joint_xyz"], dtype=np.float64)
            if j == 0:
                r = r - ref
            sp[jb + J_R: jb + J_R + 3] = r
            ax = np.array(ln["axis"])
            ai = int(np.argmax(np.abs(ax)))
            assert abs(abs(ax[ai]) - 1.0) < 1e-12 and np.count_nonzero(ax) == 1, "axis must be a coordinate axis"
            sp[jb + J_AXIS_IDX], sp[jb + J_AXIS_SIGN] = ai, np.sign(ax[ai])
            m, h, i_o, items = _composite(model, idx, np.zeros(3))
            sp[jb + J_M] = m
            sp[jb + J_H: jb + J_H + 3] = h
            sp[jb + J_I: jb + J_I + 6] = sym6(i_o)
            assert len(items) <= 2
            sp[jb + J_ND] = len(items)
            for t, it in enumerate(items):
                sp[jb + J_DAMP + t * DAMP_ITEM: jb + J_DAMP + (t + 1) * DAMP_ITEM] = it
            sp[jb + J_LOWER], sp[jb + J_UPPER] = ln["lower"], ln["upper"]
            sp[jb + J_HASLIM], sp[jb + J_JDAMP] = float(ln["lower"] <= ln["upper"]), ln["damping"]
        # foot sphere expressed in the shank (joint 3) frame
        fidx = _link_index(model, "link_%s4" % leg)
        shank = _link_index(model, "link_%s3" % leg)
        r_f, t_f = _pose_in_ancestor(model, fidx, shank)
        col = links[fidx]["collisions"][0]
        assert col["type"] == "sphere"
        center = r_f @ np.array(col["xyz"]) + t_f
        sp[lb + L_FOOT: lb + L_FOOT + 6] = [*center, col["radius"], model["foot_friction"], fidx]
    return blob


# --------------------------------------------------------------------------- IO
_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_MODEL_JSON = os.path.join(_HERE, "max_model.json")


def save_model(model, path=DEFAULT_MODEL_JSON):
    with open(path, "w") as f:
        json.dump(model, f, indent=1)


def load_model(path=DEFAULT_MODEL_JSON):
    with open(path) as f:
        return json.load(f)


def load_model_blob(path=DEFAULT_MODEL_JSON, foot_friction=None):
    model = load_model(path)
    if foot_friction is not None:
        model["foot_friction"] = float(foot_friction)
    return pack_model(model)


def write_layout_header(path):
    """Emit include/llq_model_layout.h from the constants above (single source of truth)."""
    names = ["LLQ_MODEL_MAGIC", "HDR", "H_MAGIC", "H_VERSION", "H_NLINKS", "H_NDOF", "H_OFF_GENERIC",
             "H_OFF_SPHERES", "H_NSPHERES", "H_OFF_SPECIAL", "H_TOTAL", "H_OFF_PROXIES", "H_NPROXIES", "PROXY", "GL", "G_PARENT", "G_JTYPE", "G_DOF",
             "G_JXYZ", "G_JROT", "G_AXIS", "G_MASS", "G_COM", "G_RIN", "G_IDIAG", "G_LOWER", "G_UPPER",
             "G_JDAMP", "G_HASLIM", "G_ICLINK", "SPH", "S_QI", "S_BASE_M", "S_BASE_H", "S_BASE_I", "S_BASE_ND",
             "S_BASE_DAMP", "DAMP_ITEM", "S_LEGS", "LJ", "J_R", "J_AXIS_IDX", "J_AXIS_SIGN", "J_M", "J_H", "J_I",
             "J_ND", "J_DAMP", "J_LOWER", "J_UPPER", "J_HASLIM", "J_JDAMP", "LEG", "L_FOOT", "S_TOTAL"]
    g = globals()
    lines = ["/* GENERATED by lifelike_agility_and_play_b200/model/compile_model.py -- do not edit.",
             " * Layout of the float64 robot-model blob passed to llq_load_model() (include/llq.h).",
             " * Generic section: any kinematic tree (used by the CPU oracle).",
             " * Special section: floating base + 4 legs x 3 revolute joints (used by the CUDA engine).",
             " * Source numbers: reference max.urdf (legged_robot/data/urdf/max.urdf:1-752) as imported by",
             " * PyBullet (legged_robot.py:207-220); see compile_model.py for the importer restatement. */",
             "#ifndef LLQ_MODEL_LAYOUT_H", "#define LLQ_MODEL_LAYOUT_H"]
    for nme in names:
        key = nme if nme.startswith("LLQ_") else "LLQ_" + nme
        lines.append("#define %-20s %d" % (key, g[nme]))
    lines += ["#endif", ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("urdf")
    ap.add_argument("--out", default=DEFAULT_MODEL_JSON)
    ap.add_argument("--use-urdf-inertia", action="store_true")
    ap.add_argument("--header", default=None)
    a = ap.parse_args()
    mdl = compile_model(a.urdf, a.use_urdf_inertia)
    save_model(mdl, a.out)
    if a.header:
        write_layout_header(a.header)
    print("links", len(mdl["links"]), "dof", mdl["n_dof"], "mass", sum(l["mass"] for l in mdl["links"]))
