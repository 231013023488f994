"""Shared test helpers: independent numpy kinematics / energy / momentum of the robot model
(used to check the oracle's and the CUDA engine's dynamics against conservation laws)."""
import numpy as np
from lifelike_agility_and_play_b200.model.compile_model import load_model, pack_model, rpy_to_matrix
from lifelike_agility_and_play_b200 import _capi as capi


def quat_to_matrix(q):
    x, y, z, w = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rodrigues(axis, q):
    a = np.asarray(axis, dtype=np.float64)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * K @ K


def link_kinematics(model, state):
    """World pose / velocity of every link frame + CoM from a 37-vector state (pybullet convention)."""
    st = np.asarray(state, dtype=np.float64)
    pos, quat, lin, ang, q, qd = st[0:3], st[3:7], st[7:10], st[10:13], st[13:25], st[25:37]
    links = model["links"]
    n = len(links)
    R = [None] * n; p = [None] * n; w = [None] * n; v = [None] * n
    Rb = quat_to_matrix(quat)
    b = links[0]
    R[0] = Rb @ np.array(b["R_in"]).T
    p[0] = pos - R[0] @ np.array(b["inertial_xyz"])
    w[0] = ang.copy()
    v[0] = lin + np.cross(ang, p[0] - pos)          # velocity of the link origin
    out = []
    for i, l in enumerate(links):
        if i > 0:
            pi = l["parent_index"]
            Rj = rpy_to_matrix(l["joint_rpy"])
            Rq = np.eye(3)
            wj = np.zeros(3)
            p[i] = p[pi] + R[pi] @ np.array(l["joint_xyz"])
            if l["joint_type"] == "revolute":
                Rq = rodrigues(l["axis"], q[l["dof_index"]])
                wj = (R[pi] @ Rj @ np.array(l["axis"])) * qd[l["dof_index"]]
            R[i] = R[pi] @ Rj @ Rq
            w[i] = w[pi] + wj
            v[i] = v[pi] + np.cross(w[pi], p[i] - p[pi])
        com = p[i] + R[i] @ np.array(l["inertial_xyz"])
        vcom = v[i] + np.cross(w[i], com - p[i])
        Ic = R[i] @ np.array(l["Ic_link"]) @ R[i].T
        out.append(dict(R=R[i], p=p[i], w=w[i], v=v[i], com=com, vcom=vcom, Ic=Ic, mass=l["mass"]))
    return out


def mechanics(model, state, g=9.80665):
    ks = link_kinematics(model, state)
    M = sum(k["mass"] for k in ks)
    com = sum(k["mass"] * k["com"] for k in ks) / M
    P = sum(k["mass"] * k["vcom"] for k in ks)
    L = sum(k["Ic"] @ k["w"] + k["mass"] * np.cross(k["com"], k["vcom"]) for k in ks)   # about world origin
    KE = sum(0.5 * k["mass"] * k["vcom"] @ k["vcom"] + 0.5 * k["w"] @ k["Ic"] @ k["w"] for k in ks)
    PE = sum(k["mass"] * g * k["com"][2] for k in ks)
    return dict(mass=M, com=com, P=P, L=L, KE=KE, PE=PE, links=ks)


def foot_positions(model, state):
    ks = link_kinematics(model, state)
    names = [l["name"] for l in model["links"]]
    return np.array([ks[names.index("link_%s4" % leg)]["com"] for leg in ("FR", "FL", "HR", "HL")])


def frictionless_model_blob(model=None, joint_damping=0.0):
    model = model or load_model()
    blob = pack_model(model)
    from lifelike_agility_and_play_b200.model import compile_model as cm
    off = int(blob[cm.H_OFF_GENERIC])
    for i in range(int(blob[cm.H_NLINKS])):
        blob[off + i * cm.GL + cm.G_JDAMP] = joint_damping
    sp = int(blob[cm.H_OFF_SPECIAL])
    for k in range(4):
        for j in range(3):
            blob[sp + cm.S_LEGS + k * cm.LEG + j * cm.LJ + cm.J_JDAMP] = joint_damping
    return blob


def random_state(rng, z=5.0, vel_scale=1.0):
    st = np.zeros(37)
    st[0:3] = [rng.uniform(-1, 1), rng.uniform(-1, 1), z]
    q = rng.normal(size=4); st[3:7] = q / np.linalg.norm(q)
    st[7:10] = vel_scale * rng.normal(size=3)
    st[10:13] = vel_scale * 2 * rng.normal(size=3)
    nominal = np.array([-0.03, -0.78, 1.69] * 2 + [-0.03, -0.73, 1.57] * 2)
    st[13:25] = nominal + 0.3 * rng.normal(size=12)
    st[25:37] = vel_scale * 3 * rng.normal(size=12)
    return st
