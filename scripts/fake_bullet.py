"""A fake `pybullet` / `bullet_client` so the reference's OWN env classes can run in this container.

BUILD-CONTAINER TOOLING (used only by scripts/make_trace.py to produce tests/golden/ref_task_trace.npz).
PyBullet 2.7.3 is a third-party wheel that cannot be installed here.  This module provides the ~35
non-GUI API calls the reference hot path makes on top of the float64 oracle physics
(oracle/rbd.py + the contact model of oracle/env.py), so that the reference's unmodified
`Minitaur`, `ImitationTask`, `LocomotionGymEnv` and `WrapperEnv` code -- reward terms, termination,
reference-motion bookkeeping, observation pipeline, action filter, motor model -- produces the golden
traces.  What this pins is the reference's task/robot LOGIC; the rigid-body arithmetic under it is this
repo's own model (parity with Bullet itself stays unpinned, see oracle/__init__.py).
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import xform as xf          # noqa: E402
from oracle.rbd import UrdfTree         # noqa: E402
from oracle.env import OracleEnv, DEFAULT_PHYS   # noqa: E402

ASSETS = os.path.join(ROOT, "openroborl_b200", "assets")

DIRECT, GUI = 2, 1
VELOCITY_CONTROL, TORQUE_CONTROL, POSITION_CONTROL = 0, 1, 2
JOINT_REVOLUTE, JOINT_FIXED = 0, 4
ACTIVATION_STATE_SLEEP, ACTIVATION_STATE_ENABLE_SLEEPING, ACTIVATION_STATE_DISABLE_WAKEUP = 2, 1, 32
URDF_USE_SELF_COLLISION = 8
COV_ENABLE_RENDERING, COV_ENABLE_GUI, COV_ENABLE_SINGLE_STEP_RENDERING = 7, 1, 13


class _Body(object):
    def __init__(self, robot_name, urdf, pos, orn, fixed):
        self.robot_name = robot_name
        self.tree = UrdfTree(urdf)
        self.fixed = fixed
        self.pos = np.array(pos, dtype=np.float64)
        self.orn = np.array(orn, dtype=np.float64)
        self.lin = np.zeros(3); self.ang = np.zeros(3)
        n = self.tree.n_links
        self.q = np.zeros(n); self.qd = np.zeros(n)
        self.torque = np.zeros(n)
        self.phys = None      # OracleEnv used as the physics engine of a dynamic body


class BulletClient(object):
    """pybullet_utils.bullet_client.BulletClient(connection_mode=...) look-alike."""

    def __init__(self, connection_mode=DIRECT):
        self._bodies = {}
        self._next = 0
        self._search = ASSETS
        self._dt = 1.0 / 240.0
        self._gravity = (0.0, 0.0, -9.8)
        for k, v in globals().items():
            if k.isupper() and isinstance(v, int):
                setattr(self, k, v)

    # ---- world ------------------------------------------------------------------------------------
    def setAdditionalSearchPath(self, path):
        pass   # the bundled synthetic fixtures are always used (pybullet_data is not available)

    def resetSimulation(self):
        self._bodies, self._next = {}, 0

    def setPhysicsEngineParameter(self, **kw):
        pass

    def setTimeStep(self, dt):
        self._dt = dt

    def setGravity(self, x, y, z):
        self._gravity = (x, y, z)

    def configureDebugVisualizer(self, *a, **k):
        pass

    def loadURDF(self, path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), useFixedBase=False, flags=0):
        uid = self._next
        self._next += 1
        if "plane" in path:
            self._bodies[uid] = None
            return uid
        robot = "laikago" if "laikago" in path else "mini_cheetah"
        self._bodies[uid] = _Body(robot, os.path.join(self._search, path), basePosition, baseOrientation, useFixedBase)
        return uid

    # ---- joints ------------------------------------------------------------------------------------
    def getNumJoints(self, uid):
        return self._bodies[uid].tree.n_links

    def getJointInfo(self, uid, j):
        t = self._bodies[uid].tree
        mov = bool(t.movable[j])
        qi = 7 + int(t.qidx[j]) if mov else -1
        ui = 6 + int(t.qidx[j]) if mov else -1
        return (j, t.joint_names[j].encode("UTF-8"), JOINT_REVOLUTE if mov else JOINT_FIXED, qi, ui, 0, 0.0, 0.0,
                float(t.lower[j]), float(t.upper[j]), 0.0, 0.0, t.link_names[j].encode("UTF-8"), tuple(t.axis[j]),
                (0, 0, 0), (0, 0, 0, 1), int(t.parent[j]))

    def getJointState(self, uid, j):
        b = self._bodies[uid]
        return (float(b.q[j]), float(b.qd[j]), (0.0,) * 6, float(b.torque[j]))

    def getJointStates(self, uid, ids):
        return [self.getJointState(uid, j) for j in ids]

    def getJointStateMultiDof(self, uid, j):
        b = self._bodies[uid]
        if b.tree.movable[j]:
            return ((float(b.q[j]),), (float(b.qd[j]),), (0.0,) * 6, (float(b.torque[j]),))
        return ((), (), (0.0,) * 6, ())

    def resetJointState(self, uid, j, targetValue, targetVelocity=0):
        b = self._bodies[uid]
        b.q[j], b.qd[j] = targetValue, targetVelocity

    def resetJointStateMultiDof(self, uid, j, targetValue, targetVelocity=None):
        b = self._bodies[uid]
        b.q[j] = targetValue[0]
        b.qd[j] = targetVelocity[0] if targetVelocity is not None else 0.0

    def setJointMotorControl2(self, bodyIndex, jointIndex, controlMode, targetVelocity=0, force=0, **kw):
        if controlMode == TORQUE_CONTROL:
            self._bodies[bodyIndex].torque[jointIndex] = force
        elif controlMode == VELOCITY_CONTROL and force != 0:
            raise NotImplementedError("velocity motors with non-zero force (joint friction randomisation)")

    def setJointMotorControlArray(self, bodyIndex, jointIndices, controlMode, forces=None, **kw):
        assert controlMode == TORQUE_CONTROL
        for j, f in zip(jointIndices, forces):
            self._bodies[bodyIndex].torque[j] = f

    # ---- base ----------------------------------------------------------------------------------------
    def resetBasePositionAndOrientation(self, uid, pos, orn):
        b = self._bodies[uid]
        b.pos, b.orn = np.array(pos, dtype=np.float64), np.array(orn, dtype=np.float64)

    def resetBaseVelocity(self, uid, lin, ang):
        b = self._bodies[uid]
        b.lin, b.ang = np.array(lin, dtype=np.float64), np.array(ang, dtype=np.float64)

    def getBasePositionAndOrientation(self, uid):
        b = self._bodies[uid]
        return tuple(b.pos), tuple(b.orn)

    def getBaseVelocity(self, uid):
        b = self._bodies[uid]
        return tuple(b.lin), tuple(b.ang)

    def _qvec(self, b):
        return b.q[b.tree.movable]

    def getLinkState(self, uid, j, **kw):
        b = self._bodies[uid]
        com = b.tree.link_com(b.pos[None], b.orn[None], self._qvec(b)[None])[0, j]
        return (tuple(com), (0, 0, 0, 1), (0, 0, 0), (0, 0, 0, 1), tuple(com), (0, 0, 0, 1))

    def getDynamicsInfo(self, uid, link):
        t = self._bodies[uid].tree
        return (float(t.mass[link + 1]), 0.5, tuple(np.diag(t.I_link[link + 1])))

    def changeDynamics(self, *a, **k):
        if "mass" in k or "localInertiaDiagonal" in k or "lateralFriction" in k:
            raise NotImplementedError("randomizer setters are not needed for the test-mode traces")

    def setCollisionFilterGroupMask(self, *a, **k):
        pass

    def changeVisualShape(self, *a, **k):
        pass

    def createConstraint(self, *a, **k):
        raise NotImplementedError

    # ---- transforms ------------------------------------------------------------------------------------
    def invertTransform(self, position, orientation):
        qi = xf.quaternion_inverse(np.array(orientation, dtype=np.float64))
        return tuple(-xf.QuaternionRotatePoint(np.array(position, dtype=np.float64), qi)), tuple(qi)

    def multiplyTransforms(self, positionA, orientationA, positionB, orientationB):
        qa, qb = np.array(orientationA, dtype=np.float64), np.array(orientationB, dtype=np.float64)
        pos = np.array(positionA, dtype=np.float64) + xf.QuaternionRotatePoint(np.array(positionB, dtype=np.float64), qa)
        return tuple(pos), tuple(xf.quaternion_multiply(qa, qb))

    def getEulerFromQuaternion(self, q):
        return tuple(xf.bullet_euler_from_quaternion(np.array(q, dtype=np.float64)))

    def getQuaternionFromEuler(self, rpy):
        return tuple(xf.bullet_quaternion_from_euler(np.array(rpy, dtype=np.float64)))

    # ---- contacts + stepping -----------------------------------------------------------------------------
    def _engine(self, b):
        if b.phys is None:
            clip = dict(wrap=True, frame_duration=1.0, cycle_pos=False, cycle_rot=False,
                        frames=np.tile(np.array([0, 0, 0, 0, 0, 0, 1.0] + [0.0] * 12), (2, 1)))
            urdf = "laikago/laikago_toes_limits.urdf" if b.robot_name == "laikago" else "mini_cheetah/mini_cheetah.urdf"
            b.phys = OracleEnv(b.robot_name, os.path.join(ASSETS, urdf), [clip], 1, phys=dict(DEFAULT_PHYS, gravity=self._gravity))
        return b.phys

    def getContactPoints(self, bodyA=None, bodyB=None):
        """Every collision point of bodyA within its contact margin of the plane; tuple[3] = link index."""
        b = self._bodies[bodyA]
        e = self._engine(b)
        k = b.tree.fk(b.pos[None], b.orn[None], self._qvec(b)[None])
        out = []
        for c in range(len(e.cp_body)):
            body = e.cp_body[c]
            P = k["x"][0, body] + k["R"][0, body] @ e.cp_local[c]
            if P[2] - e.cp_r[c] - e.cp_margin[c] < 0:
                out.append((0, bodyA, bodyB, int(body) - 1, -1, tuple(P), tuple(P), (0, 0, 1), float(P[2] - e.cp_r[c]), 0.0))
        return out

    def stepSimulation(self):
        for uid, b in self._bodies.items():
            if b is None or b.fixed:
                continue
            e = self._engine(b)
            assert abs(e.spec.time_step - self._dt) < 1e-15
            e.pos, e.quat, e.q = b.pos[None].copy(), b.orn[None].copy(), self._qvec(b)[None].copy()
            e.lin, e.ang, e.qd = b.lin[None].copy(), b.ang[None].copy(), b.qd[b.tree.movable][None].copy()
            tau_motor_space = np.zeros((1, 12))
            tau_urdf = b.torque[b.tree.movable][None]
            # joint_torques() expects motor-space torques; feed URDF-space ones through the inverse mapping
            tau_motor_space[:, :] = (tau_urdf[:, e.motor_q]) * e.spec.joint_dir
            nudot = e.forward_dynamics(e.joint_torques(tau_motor_space))
            e.integrate(nudot)
            b.pos, b.orn = e.pos[0].copy(), e.quat[0].copy()
            b.lin, b.ang = e.lin[0].copy(), e.ang[0].copy()
            b.q[b.tree.movable] = e.q[0]
            b.qd[b.tree.movable] = e.qd[0]
            b.torque[:] = 0.0          # Bullet clears applied joint torques every step


def install(reference_root):
    """Register the stub modules and put the reference tree on sys.path."""
    import collections
    import collections.abc
    collections.Sequence = collections.abc.Sequence            # minitaur.py:170 (removed in Python 3.10)
    me = sys.modules[__name__]
    pyb = types.ModuleType("pybullet")
    for k, v in vars(me).items():
        if k.isupper() and isinstance(v, int):
            setattr(pyb, k, v)
    pyb.configureDebugVisualizer = lambda *a, **k: None
    pyb.addUserDebugParameter = lambda *a, **k: 0
    sys.modules["pybullet"] = pyb
    bc = types.ModuleType("pybullet_utils.bullet_client")
    bc.BulletClient = BulletClient
    pu = sys.modules.get("pybullet_utils") or types.ModuleType("pybullet_utils")
    pu.bullet_client = bc
    sys.modules["pybullet_utils"] = pu
    sys.modules["pybullet_utils.bullet_client"] = bc
    pd = types.ModuleType("pybullet_data")
    pd.getDataPath = lambda: ASSETS
    sys.modules["pybullet_data"] = pd
    # gym: spaces.Box / spaces.Dict / Env / utils.seeding
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
            self.shape, self.dtype = self.low.shape, np.dtype(dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

    class Dict(object):
        def __init__(self, spaces_):
            import collections as c
            self.spaces = c.OrderedDict(sorted(spaces_.items()))     # gym.spaces.Dict sorts plain dicts by key

    spaces.Box, spaces.Dict = Box, Dict
    gym.spaces = spaces
    gym.Env = object
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    utils.seeding = seeding
    gym.utils = utils
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding})
    tf = types.ModuleType("tensorflow")
    tf.logging = types.SimpleNamespace(info=lambda *a, **k: None)
    sys.modules["tensorflow"] = tf
    sys.path.insert(0, os.path.join(reference_root, "OpenRoboRL"))
