"""URDF -> compiled quadruped model (host side of the drop-in boundary).

Replaces what the reference gets from `pybullet.loadURDF` plus the id bookkeeping of
`Minitaur._build_urdf_ids/_build_motor_id_list` (reference OpenRoboRL/envs/quadruped_robot/robots/
minitaur.py:804-864, 891-903): the kinematic tree is parsed, link ids are classified with the reference's
joint-name regexes, fixed child links are merged into their parent body for dynamics (but kept as frames
for end-effector positions), every body frame is re-oriented so its joint axis is the local z axis, and
collision primitives are turned into a set of spheres / corner points for the plane contact model.

PyBullet conventions reproduced: link index == joint index, depth-first in file order; the base state is
the pose of the base *inertial* frame (COM); getLinkState()[0] is the COM of a link.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import _abi
from . import robots as robots_pkg

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def _floats(text, default):
    return np.array([float(v) for v in text.split()], dtype=np.float64) if text else np.array(default, dtype=np.float64)


def _rpy(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _origin(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    return _rpy(_floats(o.get("rpy"), [0, 0, 0])), _floats(o.get("xyz"), [0, 0, 0])


def _align_z(axis):
    """Rotation A with A @ e_z = axis (unit)."""
    a = axis / np.linalg.norm(axis)
    z = np.array([0.0, 0.0, 1.0])
    c = float(a @ z)
    if c > 1 - 1e-12:
        return np.eye(3)
    if c < -1 + 1e-12:
        return np.diag([1.0, -1.0, -1.0])
    v = np.cross(z, a)
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + K + K @ K / (1 + c)


class UrdfLink(object):
    def __init__(self, elem):
        self.name = elem.get("name")
        inertial = elem.find("inertial")
        if inertial is None:
            self.mass, self.R_in, self.com, self.I = 0.0, np.eye(3), np.zeros(3), np.zeros((3, 3))
        else:
            self.R_in, self.com = _origin(inertial)
            self.mass = float(inertial.find("mass").get("value"))
            it = inertial.find("inertia")
            g = lambda k: float(it.get(k, 0.0))
            self.I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
        self.collisions = []
        for c in elem.findall("collision"):
            R, p = _origin(c)
            geom = c.find("geometry")
            if geom is None:
                continue
            if geom.find("sphere") is not None:
                self.collisions.append(("sphere", [float(geom.find("sphere").get("radius"))], R, p))
            elif geom.find("box") is not None:
                self.collisions.append(("box", list(_floats(geom.find("box").get("size"), [0, 0, 0])), R, p))
            elif geom.find("cylinder") is not None:
                e = geom.find("cylinder")
                self.collisions.append(("cylinder", [float(e.get("radius")), float(e.get("length"))], R, p))
            elif geom.find("capsule") is not None:
                e = geom.find("capsule")
                self.collisions.append(("capsule", [float(e.get("radius")), float(e.get("length"))], R, p))
            # meshes have no closed-form plane contact here; they are skipped (documented in DESIGN.md)
        self.lateral_friction = 0.5      # Bullet's default when the URDF gives none
        ct = elem.find("contact")
        if ct is not None and ct.find("lateral_friction") is not None:
            self.lateral_friction = float(ct.find("lateral_friction").get("value"))


class UrdfJoint(object):
    def __init__(self, elem):
        self.name = elem.get("name")
        self.type = elem.get("type")
        self.parent = elem.find("parent").get("link")
        self.child = elem.find("child").get("link")
        self.R, self.p = _origin(elem)
        ax = elem.find("axis")
        self.axis = _floats(ax.get("xyz") if ax is not None else None, [1, 0, 0])
        lim = elem.find("limit")
        self.lower, self.upper = -np.inf, np.inf
        if self.type == "revolute" and lim is not None:
            self.lower, self.upper = float(lim.get("lower", 0.0)), float(lim.get("upper", 0.0))
        dyn = elem.find("dynamics")
        self.damping = float(dyn.get("damping", 0.0)) if dyn is not None else 0.0


def parse_urdf(path):
    root = ET.parse(path).getroot()
    links = {l.get("name"): UrdfLink(l) for l in root.findall("link")}
    joints = [UrdfJoint(j) for j in root.findall("joint")]
    children = {j.child for j in joints}
    roots = [n for n in links if n not in children]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link, found %r" % roots)
    ordered = []   # (joint, parent link index)

    def visit(name, idx):
        for j in joints:
            if j.parent == name:
                me = len(ordered)
                ordered.append((j, idx))
                visit(j.child, me)
    visit(roots[0], -1)
    return links, roots[0], ordered


def find_urdf(urdf_file, search_path=None):
    """Resolve a pybullet_data-relative URDF name: user search path first, then the bundled fixtures."""
    cands = []
    if search_path:
        cands.append(os.path.join(search_path, urdf_file))
    cands.append(os.path.join(ASSET_DIR, urdf_file))
    cands.append(urdf_file)
    for c in cands:
        if os.path.exists(c):
            return c
    raise FileNotFoundError("URDF %r not found in %r" % (urdf_file, cands))


def _shape_points(kind, prm, R, p):
    """Sphere-set stand-in of a collision primitive in the link frame: [(point, radius)], bounding radius."""
    if kind == "sphere":
        return [(p.copy(), prm[0])], prm[0]
    if kind == "box":
        h = np.array(prm) / 2.0
        pts = [(p + R @ (h * np.array([sx, sy, sz])), 0.0) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
        return pts, float(np.linalg.norm(h))
    if kind in ("cylinder", "capsule"):
        r, l = prm
        pts = [(p + R @ np.array([0, 0, s * l / 2.0]), r) for s in (-1, 1)]
        return pts, float(np.hypot(r, l / 2.0) if kind == "cylinder" else r + l / 2.0)
    raise ValueError(kind)


class RobotModel(object):
    """Compiled model + the reference's id bookkeeping.  `.c` is the ctypes struct handed to orl_create."""

    def __init__(self, name_robot, urdf_search_path=None, mu_plane=1.0, margin_factor=0.02):
        spec = robots_pkg.get(name_robot)
        self.spec = spec
        self.name = name_robot
        self.urdf_path = find_urdf(spec.URDF_FILENAME, urdf_search_path)
        links, base_name, ordered = parse_urdf(self.urdf_path)
        self.joint_names = [j.name for j, _ in ordered]
        self.link_names = [j.child for j, _ in ordered]
        self.num_joints = len(ordered)
        self.joint_name_to_id = {n: i for i, n in enumerate(self.joint_names)}        # minitaur.py:804-810
        for j, _ in ordered:
            if j.type not in ("revolute", "continuous", "fixed"):
                raise ValueError("Only support 1D and 0D joints at the moment.")      # imitation_task.py:871-879
        # ---- link id classification, minitaur.py:812-851 -------------------------------------
        chassis, motor, knee, foot = [-1], [], [], []
        for i, n in enumerate(self.joint_names):
            if spec.PATTERN[0].match(n):
                chassis.append(i)
            elif spec.PATTERN[1].match(n):
                motor.append(i)
            elif spec.PATTERN[2].match(n):
                knee.append(i)
            elif spec.PATTERN[3].match(n):
                foot.append(i)
            else:
                raise ValueError("Unknown category of joint %s" % n)
        self.chassis_link_ids = sorted(chassis)
        self.motor_link_ids = sorted(motor)
        self.knee_link_ids = list(knee)
        self.leg_link_ids = sorted(knee + foot)
        self.foot_link_ids = sorted(foot + knee)
        self.motor_id_list = [self.joint_name_to_id[n] for n in spec.MOTOR_NAMES]     # minitaur.py:860-864
        # ---- topology: 4 legs x 3 revolute joints (+ fixed children) ------------------------
        parent = [p for _, p in ordered]
        movable = [j.type != "fixed" for j, _ in ordered]
        leg_roots = [i for i in range(self.num_joints) if parent[i] == -1 and movable[i]]
        if len(leg_roots) != _abi.NUM_LEGS or sum(movable) != _abi.NUM_MOTORS:
            raise ValueError("expected a floating base with 4 legs of 3 revolute joints, got %d legs / %d dofs"
                             % (len(leg_roots), sum(movable)))
        if any(parent[i] == -1 and not movable[i] for i in range(self.num_joints)):
            raise ValueError("fixed links attached to the base are not supported")
        chains = []
        for r in leg_roots:
            chain = [r]
            while True:
                kids = [i for i in range(self.num_joints) if parent[i] == chain[-1] and movable[i]]
                if not kids:
                    break
                if len(kids) > 1:
                    raise ValueError("branching leg")
                chain.append(kids[0])
            if len(chain) != _abi.LEG_DOF:
                raise ValueError("leg with %d revolute joints" % len(chain))
            chains.append(chain)
        rev = [i for i in range(self.num_joints) if movable[i]]
        self.q_index = {i: n for n, i in enumerate(rev)}            # joint id -> index into pose[7:]
        for l, chain in enumerate(chains):
            for k, i in enumerate(chain):
                assert self.q_index[i] == 3 * l + k, "revolute joints of a leg must be consecutive in URDF order"
        self.leg_joint_ids = chains
        motor_of_joint = {jid: m for m, jid in enumerate(self.motor_id_list)}
        # ---- compile ------------------------------------------------------------------------
        c = _abi.OrlRobotModel()
        base = links[base_name]
        R_lb = base.R_in                       # base link frame <- base state (inertial) frame
        c.base_mass = base.mass
        Ib = base.I                            # already expressed in the inertial frame
        c.base_inertia[:] = [Ib[0, 0], Ib[0, 1], Ib[0, 2], Ib[1, 1], Ib[1, 2], Ib[2, 2]]
        c.init_pos[:] = [float(v) for v in spec.INIT_POSITION]
        c.init_quat[:] = [float(v) for v in spec.INIT_QUAT]
        c.max_angle_change = spec.MAX_MOTOR_ANGLE_CHANGE_PER_STEP
        c.n_links = self.num_joints
        if self.num_joints > _abi.MAX_LINKS:
            raise ValueError("too many links")
        # base -> link frame transform expressed in the state frame: x_state = R_lb^T (x_link - com)
        base_pts = self._collision_points(base, lambda x: R_lb.T @ (x - base.com), -1, False, mu_plane, margin_factor,
                                          com_state=np.zeros(3))
        leg_cps = [[] for _ in range(_abi.NUM_LEGS)]
        self.mass_total = base.mass
        for l, chain in enumerate(chains):
            L = c.leg[l]
            A_parent, parent_is_base = None, True
            for k, jid in enumerate(chain):
                joint = ordered[jid][0]
                link = links[joint.child]
                A = _align_z(joint.axis)                                   # link frame -> body frame alignment
                if parent_is_base:
                    Rp, pp = R_lb.T, -R_lb.T @ base.com                   # state frame <- base link frame
                else:
                    Rp, pp = A_parent.T, np.zeros(3)
                R_fix = Rp @ joint.R @ A
                p_fix = Rp @ joint.p + pp
                for r in range(3):
                    for cc in range(3):
                        L.R_fix[k][3 * r + cc] = R_fix[r, cc]
                    L.p_fix[k][r] = p_fix[r]
                to_body = (lambda AA: (lambda x: AA.T @ x))(A)
                # bodies rigidly attached below this joint (fixed children, recursively)
                parts = [(link, np.eye(3), np.zeros(3), jid)]
                stack = [(jid, np.eye(3), np.zeros(3))]
                while stack:
                    pid, Rw, pw = stack.pop()
                    for i in range(self.num_joints):
                        if parent[i] == pid and not movable[i]:
                            jf = ordered[i][0]
                            Rc, pc = Rw @ jf.R, pw + Rw @ jf.p
                            parts.append((links[jf.child], Rc, pc, i))
                            stack.append((i, Rc, pc))
                m_tot = sum(pt[0].mass for pt in parts)
                if m_tot <= 0:
                    raise ValueError("massless body on leg %d" % l)
                coms = [pc + Rc @ pl.com for pl, Rc, pc, _ in parts]          # in this joint's LINK frame
                com = sum(pl.mass * cm for (pl, _, _, _), cm in zip(parts, coms)) / m_tot
                I = np.zeros((3, 3))
                I_par = np.zeros((3, 3))        # parallel-axis share: scales with the randomised mass ratio
                for (pl, Rc, pc, _), cm in zip(parts, coms):
                    Rl = Rc @ pl.R_in
                    d = cm - com
                    I_par += pl.mass * ((d @ d) * np.eye(3) - np.outer(d, d))
                    I += Rl @ pl.I @ Rl.T
                I = I + I_par
                com_b, I_b, I_pb = to_body(com), A.T @ I @ A, A.T @ I_par @ A
                # randomizer bookkeeping (minitaur.py:935-1070): class of the merged links, knee-friction sample
                classes = {0 if lid in self.chassis_link_ids else 1 for _, _, _, lid in parts}
                L.rand_class[k] = classes.pop() if len(classes) == 1 else -1
                L.friction_draw[k] = self.foot_link_ids.index(jid) if jid in self.foot_link_ids else -1
                L.inertia_par[k][:] = [I_pb[0, 0], I_pb[0, 1], I_pb[0, 2], I_pb[1, 1], I_pb[1, 2], I_pb[2, 2]]
                L.mass[k] = m_tot
                self.mass_total += m_tot
                L.com[k][:] = list(com_b)
                L.inertia[k][:] = [I_b[0, 0], I_b[0, 1], I_b[0, 2], I_b[1, 1], I_b[1, 2], I_b[2, 2]]
                big = 1e30
                L.lower[k] = max(joint.lower, -big)
                L.upper[k] = min(joint.upper, big)
                L.damping[k] = joint.damping
                m = motor_of_joint[jid]
                L.motor[k] = m
                L.dir[k] = float(spec.JOINT_DIRECTIONS[m]); L.offset[k] = float(spec.JOINT_OFFSETS[m])
                L.kp[k] = float(spec.motor_kp[m]); L.kd[k] = float(spec.motor_kd[m])
                L.init_angle[k] = float(spec.INIT_MOTOR_ANGLES[m])
                for pl, Rc, pc, lid in parts:
                    cm = to_body(pc + Rc @ pl.com)
                    c.link_leg[lid] = l; c.link_body[lid] = k
                    c.link_com[lid][:] = list(cm)
                    is_foot = lid in self.foot_link_ids
                    if is_foot:
                        if L.n_ee >= 2:
                            raise ValueError("more than 2 end-effector links on a leg")
                        L.ee_body[L.n_ee] = k
                        L.ee_local[L.n_ee][:] = list(cm)
                        L.n_ee += 1
                    tf = (lambda Rc_, pc_, AA: (lambda x: AA.T @ (pc_ + Rc_ @ x)))(Rc, pc, A)
                    leg_cps[l] += self._collision_points(pl, tf, k, is_foot, mu_plane, margin_factor, com_state=cm)
                A_parent, parent_is_base = A, False
        # distribute the base's points round-robin over the four legs' lists
        for i, cp in enumerate(base_pts):
            leg_cps[i % _abi.NUM_LEGS].append(cp)
        for l in range(_abi.NUM_LEGS):
            if len(leg_cps[l]) > _abi.MAX_CP:
                raise ValueError("too many collision points on leg %d (%d > %d)" % (l, len(leg_cps[l]), _abi.MAX_CP))
            c.leg[l].n_cp = len(leg_cps[l])
            for i, (loc, rad, mu, margin, body, is_foot) in enumerate(leg_cps[l]):
                cp = c.leg[l].cp[i]
                cp.local[:] = list(loc); cp.radius = rad; cp.mu = mu; cp.margin = margin
                cp.body = body; cp.is_foot = int(is_foot)
        self.c = c

    @staticmethod
    def _collision_points(link, to_body, body, is_foot, mu_plane, margin_factor, com_state):
        out = []
        for kind, prm, R, p in link.collisions:
            pts, brad = _shape_points(kind, prm, R, p)
            # Bullet's contact-breaking threshold ~ 0.02 * angular-motion disc of the shape about the
            # link COM (unpinned)
            disc = brad + float(np.linalg.norm(p - link.com))
            for x, r in pts:
                out.append((to_body(x), float(r), link.lateral_friction * mu_plane, margin_factor * disc, body, is_foot))
        return out

    # defaults used by Minitaur.reset_pose / ImitationTask._record_default_pose
    def default_joint_pose(self):
        """(INIT + OFFSET) * DIR in URDF joint order (minitaur.py:1190-1194)."""
        s = self.spec
        mot = (np.asarray(s.INIT_MOTOR_ANGLES) + np.asarray(s.JOINT_OFFSETS)) * np.asarray(s.JOINT_DIRECTIONS)
        out = np.zeros(12)
        for m, jid in enumerate(self.motor_id_list):
            out[self.q_index[jid]] = mot[m]
        return out
