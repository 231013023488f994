"""Model compiler: numbers derived from max.urdf."""
import os
import re

import numpy as np

from lifelike_agility_and_play_b200.model import compile_model as cm
from helpers import foot_positions, mechanics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_total_mass_and_link_order(model):
    links = model["links"]
    assert len(links) == 23 and model["n_dof"] == 12
    assert abs(sum(l["mass"] for l in links) - 13.00021) < 1e-5
    names = [l["name"] for l in links]
    # pybullet link index = position among non-base links (URDF_MAINTAIN_LINK_ORDER): LR:225-231 rely on these
    idx = {n: i - 1 for i, n in enumerate(names)}
    assert [idx["link_%s%d" % (g, j)] for g in ("FR", "FL", "HR", "HL") for j in (1, 2, 3)] == [0, 1, 2, 5, 6, 7, 10, 11, 12, 15, 16, 17]
    assert [idx["link_%s4" % g] for g in ("FR", "FL", "HR", "HL")] == [3, 8, 13, 18]
    assert [idx["link_%sW" % g] for g in ("FR", "FL", "HR", "HL")] == [4, 9, 14, 19]
    assert [idx["link_front_handle"], idx["link_hind_handle"]] == [20, 21]
    rev = [l for l in links if l["joint_type"] == "revolute"]
    assert [l["dof_index"] for l in rev] == list(range(12))


def test_principal_axes_rotation_angles(model):
    """Bullet's Jacobi diagonalisation of the URDF tensors (body 2.1 deg, hips 5.7, thighs 9.6-9.8, shanks 2.5-2.7)."""
    def ang(name):
        R = np.array(next(l for l in model["links"] if l["name"] == name)["R_in"])
        assert abs(np.linalg.det(R) - 1) < 1e-9 and np.allclose(R @ R.T, np.eye(3), atol=1e-9)
        return np.degrees(np.arccos((np.trace(R) - 1) / 2))
    assert abs(ang("body") - 2.11) < 0.05
    assert abs(ang("link_FR1") - 5.7) < 0.1
    assert 9.5 < ang("link_FR2") < 9.9 and 9.5 < ang("link_HL2") < 9.9
    assert 2.4 < ang("link_FR3") < 2.8 and 2.4 < ang("link_HR3") < 2.8


def test_bullet_diagonalize_reconstructs():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = rng.normal(size=(3, 3)); m = a @ a.T + np.eye(3)
        d, rot = cm.bullet_diagonalize(m)
        assert np.allclose(rot @ np.diag(d) @ rot.T, m, atol=1e-5 * np.abs(m).max())
        assert abs(np.linalg.det(rot) - 1) < 1e-9


def test_inertia_from_collision_aabb_default(model):
    """Without URDF_USE_INERTIA_FROM_FILE (LR:212-217) Bullet replaces the URDF moments by the collision-AABB box inertia."""
    assert model["use_urdf_inertia"] is False
    body = model["links"][0]
    d = np.array(body["inertia_diag"])
    assert np.all(d > 0) and not np.allclose(d, np.diag(np.array(body["inertia_urdf"])), rtol=0.2)
    # a box 0.283 x 0.205 x 0.11 seen from a frame tilted by 2.1 deg: slightly larger than the exact box inertia
    exact = body["mass"] / 12 * np.array([0.205 ** 2 + 0.11 ** 2, 0.283 ** 2 + 0.11 ** 2, 0.283 ** 2 + 0.205 ** 2])
    assert np.all(d >= exact * 0.999) and np.all(d < exact * 1.25)
    foot = next(l for l in model["links"] if l["name"] == "link_FR4")
    assert foot["mass"] == 0 and np.allclose(foot["inertia_diag"], 0)


def test_zero_pose_feet(model):
    """K5: zero-pose feet at (+-0.195, -+0.15, -0.4515) in the base *link* frame."""
    st = np.zeros(37); st[6] = 1.0
    body = model["links"][0]
    R_I = np.array(body["R_in"])
    # state speaks in the base inertial frame: put the link frame at the origin with identity orientation
    from lifelike_agility_and_play_b200.model.compile_model import matrix_to_quat_xyzw
    st[3:7] = matrix_to_quat_xyzw(R_I)
    st[0:3] = np.array(body["inertial_xyz"])
    f = foot_positions(model, st)
    want = np.array([[0.195, -0.15, -0.4515], [0.195, 0.15, -0.4515], [-0.195, -0.15, -0.4515], [-0.195, 0.15, -0.4515]])
    assert np.allclose(f, want, atol=1e-9)


def test_special_section_matches_generic(model, blob):
    """The CUDA engine's folded (composite) tables must describe the same mass distribution as the generic tree."""
    off = int(blob[cm.H_OFF_SPECIAL])
    sp = blob[off:]
    m_tot = sp[cm.S_BASE_M] + sum(sp[cm.S_LEGS + k * cm.LEG + j * cm.LJ + cm.J_M] for k in range(4) for j in range(3))
    assert abs(m_tot - 13.00021) < 1e-5
    # composite CoM of the whole robot at the zero pose, from the special section vs from the generic kinematics
    st = np.zeros(37); st[6] = 1.0
    mech = mechanics(model, st)
    from helpers import quat_to_matrix
    R_I = np.array(model["links"][0]["R_in"])
    # special section works in body-link axes about the body CoM
    first = sp[cm.S_BASE_H:cm.S_BASE_H + 3].copy()
    for k in range(4):
        lb = cm.S_LEGS + k * cm.LEG
        pos = np.zeros(3)
        for j in range(3):
            jb = lb + j * cm.LJ
            pos = pos + sp[jb + cm.J_R: jb + cm.J_R + 3]       # zero pose: all link axes parallel to the body axes
            first += sp[jb + cm.J_M] * pos + sp[jb + cm.J_H: jb + cm.J_H + 3]
    com_link_axes = first / m_tot                               # relative to body CoM, link axes
    com_world = st[0:3] + R_I.T @ com_link_axes                 # state orientation = inertial frame => link axes = R_I^T
    assert np.allclose(com_world, mech["com"], atol=1e-9)


def test_layout_header_in_sync():
    txt = open(os.path.join(ROOT, "include", "llq_model_layout.h")).read()
    defs = dict(re.findall(r"#define\s+(LLQ_\w+)\s+(\d+)", txt))
    for name in ("HDR", "GL", "G_MASS", "G_RIN", "SPH", "S_QI", "S_LEGS", "LJ", "LEG", "L_FOOT", "S_TOTAL", "J_JDAMP", "DAMP_ITEM"):
        assert int(defs["LLQ_" + name]) == getattr(cm, name), name
    assert int(defs["LLQ_MODEL_MAGIC"]) == cm.LLQ_MODEL_MAGIC


def test_joint_limits_and_damping(model):
    l = {x["name"]: x for x in model["links"]}
    assert (l["link_FR1"]["lower"], l["link_FR1"]["upper"]) == (-0.872, 0.697)
    assert (l["link_FL1"]["lower"], l["link_FL1"]["upper"]) == (-0.697, 0.872)
    assert (l["link_HR3"]["lower"], l["link_HR3"]["upper"]) == (-2.3995, 2.5855)
    assert all(l["link_%s%d" % (g, j)]["damping"] == 0.1 for g in ("FR", "FL", "HR", "HL") for j in (1, 2, 3))
