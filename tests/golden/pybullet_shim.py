"""Stand-ins for the reference's missing third-party modules (pybullet, pybullet_utils, gym, tleague) so that the
*unmodified* reference classes (LeggedRobot, MotionLib, PrimitiveLevelEnv, create_tracking_game) can be imported and
executed in this container.  Only used by gen_golden_from_reference.py (which needs /root/reference) -- never at test
time, never on the GPU box.

`pybullet` is the interesting one: every call the reference makes (legged_robot.py, primitive_level_env.py) is mapped
onto oracle/libllq_cpu.so's fp64 physics through the oracle-only hooks llq_oracle_{get,set}_state64 /
llq_oracle_substep / llq_oracle_foot_positions.  What the resulting golden vectors pin is therefore the reference's
*environment logic* (call order, sub-step loop, mocap clock, interpolation, observation, reward, termination,
prioritized sampling) on top of the oracle's physics; Bullet's own numerics stay unpinned (no Bullet here)."""
import ctypes as C
import sys
import types
from collections import OrderedDict

import numpy as np

LEG_LINKS = [0, 1, 2, 5, 6, 7, 10, 11, 12, 15, 16, 17]
FOOT_LINKS = [3, 8, 13, 18]
JOINT_NAMES = []
for leg in ("FR", "FL", "HR", "HL"):
    JOINT_NAMES += ["joint_%s1" % leg, "joint_%s2" % leg, "joint_%s3" % leg, "joint_%s4" % leg, "joint_%sW" % leg]
JOINT_NAMES += ["joint_front_handle", "joint_hind_handle"]


def _proxy_table():
    from lifelike_agility_and_play_b200.model.compile_model import load_model_blob, H_OFF_PROXIES, H_NPROXIES, PROXY
    b = load_model_blob()
    o, n = int(b[H_OFF_PROXIES]), int(b[H_NPROXIES])
    t = np.asarray(b[o:o + n * PROXY], dtype=np.float64).reshape(n, PROXY)
    return t[:, 0].astype(int), t[:, 4].copy(), t[:, 5].astype(int)


PROXY_LINKS, PROXY_RADII, PROXY_KINDS = _proxy_table()      # model link (pybullet link index + 1), radius, kind


class _Body:
    def __init__(self, kind):
        self.kind = kind
        self.state = np.zeros(37)
        self.state[6] = 1.0
        self.dirty = True
        self.tau = np.zeros(12)
        self.push = None          # pending applyExternalForce (link 0, LINK_FRAME), cleared after every step
        self.foot_mu = 0.5


class FakeBulletClient:
    # constants the reference touches
    GUI, DIRECT = 1, 2
    URDF_MAINTAIN_LINK_ORDER, URDF_USE_SELF_COLLISION, URDF_ENABLE_CACHED_GRAPHICS_SHAPES = 1, 2, 4
    URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS = 8
    POSITION_CONTROL, TORQUE_CONTROL = 2, 3
    ACTIVATION_STATE_SLEEP, ACTIVATION_STATE_ENABLE_SLEEPING, ACTIVATION_STATE_DISABLE_WAKEUP = 1, 2, 4
    COV_ENABLE_RENDERING, COV_ENABLE_GUI, COV_ENABLE_SINGLE_STEP_RENDERING = 1, 2, 3
    GEOM_BOX, STATE_LOGGING_VIDEO_MP4, GEOM_CYLINDER = 3, 4, 5
    LINK_FRAME, WORLD_FRAME = 1, 2

    oracle_engine = None      # VecEngine over libllq_cpu.so with n_envs = 1 (2 for the chase-tag pair), set by the generator
    contact_breaking = 0.0005
    boxes_block_rays = False  # chase-tag generator: createMultiBody boxes (walls, flag) are seen by rays
    terrain_boxes = False     # EPMC corridor generator: static boxes are mirrored into the oracle (foot contacts) before every step
    pair_contacts = False     # chase-tag generator: robot-robot / robot-flag contact points from detection proxies
    call_log = None

    def __init__(self, connection_mode=None):
        self.bodies = []
        self.connected = True
        eng = FakeBulletClient.oracle_engine
        self._lib = eng.lib.lib
        self._h = eng._h
        for f in ("llq_oracle_get_state64", "llq_oracle_set_state64", "llq_oracle_substep", "llq_oracle_substep_push",
                  "llq_oracle_foot_positions"):
            getattr(self._lib, f).restype = C.c_int

    # ---- model loading
    def loadURDF(self, path, basePosition=None, baseOrientation=None, flags=0, globalScaling=1.0, useFixedBase=False):
        if path.endswith("max.urdf"):
            b = _Body("kinematic" if useFixedBase else "dynamic")
            if basePosition is not None:
                b.state[0:3] = basePosition
            if baseOrientation is not None:
                b.state[3:7] = baseOrientation
        else:
            b = _Body("static")
        self.bodies.append(b)
        return len(self.bodies) - 1

    def getNumJoints(self, uid):
        return 22

    def getJointInfo(self, uid, j):
        return (j, JOINT_NAMES[j].encode("utf-8"))

    def _noop(self, *a, **k):
        return None
    setCollisionFilterGroupMask = changeVisualShape = setGravity = _noop
    setPhysicsEngineParameter = setTimeStep = configureDebugVisualizer = resetDebugVisualizerCamera = _noop
    createVisualShape = _noop

    def createCollisionShape(self, shapeType=None, halfExtents=None, radius=None, height=None, **kw):
        if not hasattr(self, "shapes"):
            self.shapes = []
        if shapeType == self.GEOM_CYLINDER:          # auxiliary edge cylinder (BSE:43-104): always laid along y
            self.shapes.append(("cyl", float(radius), float(height)))
        else:
            self.shapes.append(None if halfExtents is None else np.asarray(halfExtents, dtype=np.float64))
        return len(self.shapes) - 1

    def createMultiBody(self, baseMass=0, baseCollisionShapeIndex=-1, baseVisualShapeIndex=-1, basePosition=None, baseOrientation=None, **k):
        b = _Body("static")
        shp = self.shapes[baseCollisionShapeIndex] if hasattr(self, "shapes") and baseCollisionShapeIndex >= 0 else None
        b.cyl = shp[1:] if isinstance(shp, tuple) else None
        b.box = None if isinstance(shp, tuple) else shp
        b.ray_target = FakeBulletClient.boxes_block_rays
        if basePosition is not None:
            b.state[0:3] = basePosition
        if baseOrientation is not None:
            b.state[3:7] = baseOrientation
        self.bodies.append(b)
        return len(self.bodies) - 1

    def removeBody(self, uid):
        self.bodies[uid].kind = "removed"

    def getContactPoints(self, bodyA=None, **kw):
        """Contacts of the last stepSimulation (built on that step's pre-step poses): only robot-vs-box-body pairs are modelled."""
        return list(getattr(self, "_contacts", []))

    def changeDynamics(self, uid, linkIndex=None, lateralFriction=None, **kw):
        if lateralFriction is not None and linkIndex in FOOT_LINKS and self.bodies[uid].kind == "dynamic":
            self.bodies[uid].foot_mu = float(lateralFriction)        # LR:304-308

    def applyExternalForce(self, objectUniqueId, linkIndex, forceObj, posObj, flags):
        assert linkIndex == 0 and flags == self.LINK_FRAME and not np.any(np.asarray(posObj))        # PR:73-77
        self.bodies[objectUniqueId].push = np.asarray(forceObj, dtype=np.float64).copy()

    def _static_boxes(self):
        """(uid, lo, hi) of everything a mask-6 ray can hit: the ground slab of plane.urdf (box 200 x 200 x 10 centred at
        z = -5) and every live createMultiBody box (walls, flag).  Boxes are axis aligned in all shipped arenas."""
        out = [(-2, np.array([-100.0, -100.0, -10.0]), np.array([100.0, 100.0, 0.0]))]
        for uid, o in enumerate(self.bodies):
            if o.kind == "static" and getattr(o, "box", None) is not None and np.any(o.box > 0) and getattr(o, "ray_target", False):
                c = np.asarray(o.state[0:3], dtype=np.float64)
                out.append((uid, c - o.box, c + o.box))
        return out

    def rayTestBatch(self, rayFromPositions, rayToPositions, collisionFilterMask=-1, **kw):
        """Closest hit against the static world.  Independent slab test (not the oracle's code).  Mask 6 => robots are invisible."""
        boxes = self._static_boxes()
        out = []
        for a, b in zip(np.asarray(rayFromPositions, dtype=np.float64), np.asarray(rayToPositions, dtype=np.float64)):
            d = b - a
            best = None
            for uid, lo, hi in boxes:
                t0, t1, axis_in = 0.0, 1.0, -1
                inside = bool(np.all(a > lo) and np.all(a < hi))
                hit = not inside
                if hit:
                    for ax in range(3):
                        if d[ax] == 0.0:
                            if a[ax] < lo[ax] or a[ax] > hi[ax]:
                                hit = False
                                break
                            continue
                        ta, tb = (lo[ax] - a[ax]) / d[ax], (hi[ax] - a[ax]) / d[ax]
                        if ta > tb:
                            ta, tb = tb, ta
                        if ta > t0:
                            t0, axis_in = ta, ax
                        t1 = min(t1, tb)
                        if t0 > t1:
                            hit = False
                            break
                if hit and axis_in >= 0 and (best is None or t0 < best[0]):
                    best = (t0, axis_in, uid)
            if best is not None:
                t0, axis_in, uid = best
                n = np.zeros(3); n[axis_in] = -np.sign(d[axis_in])
                out.append((uid if uid >= 0 else len(self.bodies), -1, t0, tuple(a + t0 * d), tuple(n)))
            else:
                out.append((-1, -1, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)))
        return out

    def rayTest(self, rayFromPosition, rayToPosition, collisionFilterMask=-1, **kw):
        return self.rayTestBatch([rayFromPosition], [rayToPosition], collisionFilterMask)

    def setJointMotorControlArray(self, bodyUniqueId, jointIndices, controlMode, forces=None, **kw):
        if controlMode == self.TORQUE_CONTROL:
            assert list(jointIndices) == LEG_LINKS
            self.bodies[bodyUniqueId].tau = np.asarray(forces, dtype=np.float64).copy()

    # ---- state access
    def resetBasePositionAndOrientation(self, uid, pos, orn):
        b = self.bodies[uid]
        b.state[0:3], b.state[3:7], b.dirty = pos, orn, True

    def resetBaseVelocity(self, uid, lin, ang):
        b = self.bodies[uid]
        b.state[7:10], b.state[10:13], b.dirty = lin, ang, True

    def resetJointState(self, uid, jointIndex, pos, vel=0.0):
        b = self.bodies[uid]
        d = LEG_LINKS.index(jointIndex)
        b.state[13 + d], b.state[25 + d], b.dirty = pos, vel, True

    def getBasePositionAndOrientation(self, uid):
        s = self.bodies[uid].state
        return tuple(s[0:3]), tuple(s[3:7])

    def getBaseVelocity(self, uid):
        s = self.bodies[uid].state
        return tuple(s[7:10]), tuple(s[10:13])

    def getJointStates(self, uid, indices):
        s = self.bodies[uid].state
        return [(s[13 + LEG_LINKS.index(j)], s[25 + LEG_LINKS.index(j)], (0.0,) * 6, 0.0) for j in indices]

    def _proxy_positions(self, b):
        st = np.ascontiguousarray(b.state, dtype=np.float64)
        out = np.zeros((32, 3))
        n = C.c_int32(0)
        self._lib.llq_oracle_proxy_positions.restype = C.c_int
        assert self._lib.llq_oracle_proxy_positions(self._h, st.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 32, C.byref(n)) == 0
        return out[:n.value]

    def getLinkStates(self, uid, indices, computeForwardKinematics=False, computeLinkVelocity=False):
        z3 = (0.0, 0.0, 0.0)
        b = self.bodies[uid]
        if list(indices) == FOOT_LINKS:
            st = np.ascontiguousarray(b.state, dtype=np.float64)
            out = np.zeros(12)
            rc = self._lib.llq_oracle_foot_positions(self._h, st.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
            assert rc == 0
            return [(tuple(out[3 * i:3 * i + 3]), (0, 0, 0, 1), z3, (0, 0, 0, 1), z3, (0, 0, 0, 1), z3, z3) for i in range(4)]
        # feet / wheels / handles (LR:150-156): link COM == link origin == the proxy sphere centre for these links
        pp = self._proxy_positions(b)
        by_link = {int(l) - 1: k for k, l in enumerate(PROXY_LINKS) if PROXY_KINDS[k] != 3}
        return [(tuple(pp[by_link[j]]), (0, 0, 0, 1), z3, (0, 0, 0, 1), z3, (0, 0, 0, 1), z3, z3) for j in indices]

    # ---- the physics step
    def _mirror_boxes(self, dyn):
        """EPMC corridor: hand the live static boxes (walls, hurdles, bars, cubes; not the degenerate target marker) and the
        auxiliary edge cylinders to the oracle, whose narrow phase then collides the robot's spheres with them."""
        bx = [np.r_[o.state[0:3], o.box] for o in self.bodies
              if o.kind == "static" and getattr(o, "box", None) is not None and np.any(o.box > 0)]
        arr = np.ascontiguousarray(np.array(bx, dtype=np.float64).reshape(-1, 6))
        self._lib.llq_oracle_set_boxes.restype = C.c_int
        cy = [np.r_[o.state[0:3], o.cyl] for o in self.bodies if o.kind == "static" and getattr(o, "cyl", None) is not None]
        carr = np.ascontiguousarray(np.array(cy, dtype=np.float64).reshape(-1, 5))
        self._lib.llq_oracle_set_cylinders.restype = C.c_int
        for k in range(len(dyn)):
            assert self._lib.llq_oracle_set_boxes(self._h, k, arr.ctypes.data_as(C.c_void_p), len(bx)) == 0
            assert self._lib.llq_oracle_set_cylinders(self._h, k, carr.ctypes.data_as(C.c_void_p), len(cy)) == 0

    def _narrow_phase(self):
        """Contact points for getContactPoints(), built on the pre-step poses.  PMC: robot detection proxies vs static boxes
        through the oracle's hurdle test.  Chase tag: independent numpy sphere-sphere / sphere-box tests over the model's
        proxies (link index = model link - 1), in body order."""
        self._contacts = []
        dyn = [b for b in self.bodies if b.kind == "dynamic"]
        if FakeBulletClient.pair_contacts:
            thr = FakeBulletClient.contact_breaking
            pps = [self._proxy_positions(b) for b in dyn]
            ids = [self.bodies.index(b) for b in dyn]
            for ia in range(len(dyn)):
                for ib in range(ia + 1, len(dyn)):
                    for ka in range(len(PROXY_LINKS)):
                        for kb in range(len(PROXY_LINKS)):
                            if np.linalg.norm(pps[ia][ka] - pps[ib][kb]) - PROXY_RADII[ka] - PROXY_RADII[kb] < thr:
                                self._contacts.append((0, ids[ia], ids[ib], int(PROXY_LINKS[ka]) - 1, int(PROXY_LINKS[kb]) - 1))
                for uid, o in enumerate(self.bodies):
                    if o.kind == "static" and getattr(o, "box", None) is not None and getattr(o, "is_flag", False):
                        c = np.asarray(o.state[0:3])
                        for ka in range(len(PROXY_LINKS)):
                            q = pps[ia][ka] - np.clip(pps[ia][ka], c - o.box, c + o.box)
                            if np.linalg.norm(q) - PROXY_RADII[ka] < thr:
                                self._contacts.append((0, ids[ia], uid, int(PROXY_LINKS[ka]) - 1, -1))
            return
        self._lib.llq_oracle_obstacle_hit.restype = C.c_int
        for b in dyn:
            for uid, o in enumerate(self.bodies):
                if o.kind == "static" and getattr(o, "box", None) is not None and np.any(o.box > 0):
                    x, y, z, w = o.state[3:7] / np.linalg.norm(o.state[3:7])
                    pose = np.array([o.state[0], o.state[1], np.arctan2(2 * (x * y + z * w), 1 - 2 * (y * y + z * z))])
                    st = np.ascontiguousarray(b.state, dtype=np.float64)
                    half = np.ascontiguousarray(o.box, dtype=np.float64)
                    hit = C.c_int32(0)
                    assert self._lib.llq_oracle_obstacle_hit(self._h, st.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.c_void_p),
                                                             half.ctypes.data_as(C.c_void_p), C.byref(hit)) == 0
                    if hit.value:
                        self._contacts.append((0, self.bodies.index(b), uid, -1, -1))

    def stepSimulation(self):
        dyn = [b for b in self.bodies if b.kind == "dynamic"]
        for k, b in enumerate(dyn):
            if b.dirty:
                st = np.ascontiguousarray(b.state, dtype=np.float64)
                assert self._lib.llq_oracle_set_state64(self._h, k, st.ctypes.data_as(C.c_void_p)) == 0
                b.dirty = False
        if FakeBulletClient.terrain_boxes:
            self._mirror_boxes(dyn)
        else:
            self._narrow_phase()
        for k, b in enumerate(dyn):
            tau = np.ascontiguousarray(b.tau, dtype=np.float64)
            push = None if b.push is None else np.ascontiguousarray(b.push, dtype=np.float64)
            self._lib.llq_oracle_substep_push.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double]
            assert self._lib.llq_oracle_substep_push(self._h, k, tau.ctypes.data_as(C.c_void_p),
                                                      None if push is None else push.ctypes.data_as(C.c_void_p), C.c_double(b.foot_mu)) == 0
            b.push = None                 # external forces last one step
            out = np.zeros(37)
            assert self._lib.llq_oracle_get_state64(self._h, k, out.ctypes.data_as(C.c_void_p)) == 0
            b.state = out
            b.tau = np.zeros(12)          # Bullet clears applied torques after every step

    def isConnected(self):
        return 0

    def disconnect(self):
        self.connected = False


def install():
    """Register the fake modules; returns the FakeBulletClient class."""
    from lifelike_agility_and_play_b200 import spaces as myspaces

    pb = types.ModuleType("pybullet")
    for k in dir(FakeBulletClient):
        if k.isupper():
            setattr(pb, k, getattr(FakeBulletClient, k))
    sys.modules["pybullet"] = pb
    pbu = types.ModuleType("pybullet_utils")
    bc = types.ModuleType("pybullet_utils.bullet_client")
    bc.BulletClient = FakeBulletClient
    pbu.bullet_client = bc
    sys.modules["pybullet_utils"], sys.modules["pybullet_utils.bullet_client"] = pbu, bc

    gym = types.ModuleType("gym")

    class Env:
        pass

    class Wrapper:
        def __init__(self, env):
            self.env = env

        def step(self, action):
            return self.env.step(action)

        def __getattr__(self, name):
            return getattr(self.env, name)
    gym.Env, gym.Wrapper = Env, Wrapper
    sp = types.ModuleType("gym.spaces")
    sp.Box, sp.Dict, sp.Tuple, sp.Discrete = myspaces.Box, myspaces.Dict, myspaces.Tuple, myspaces.Discrete
    gym.spaces = sp
    sys.modules["gym"], sys.modules["gym.spaces"] = gym, sp

    tl = types.ModuleType("tleague")
    tlu = types.ModuleType("tleague.utils")
    lg = types.ModuleType("tleague.utils.logger")
    lg.log = lambda *a, **k: None
    tlu.logger = lg
    tl.utils = tlu
    sys.modules["tleague"], sys.modules["tleague.utils"], sys.modules["tleague.utils.logger"] = tl, tlu, lg
    return FakeBulletClient
