"""Textbook rigid-body kinematics/dynamics on the RAW URDF tree, float64, batched.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the oracle for the arithmetic the
reference delegates to PyBullet 2.7.3 (`getLinkState`, `stepSimulation`; SURVEY.md 8a rows
a8/a14, A.9, A.10).  PARITY UNPINNED against Bullet (not installable here).  It is written to
be *independent* of the product's compiled model: no frame re-alignment, no merging of fixed
links, no spatial algebra -- just world-frame Newton-Euler per body and maximal-coordinate
Jacobians (Kane's equations), so that agreement with the CUDA CRBA/RNEA path is meaningful.

Conventions (PyBullet's): link index i == joint index i, depth-first in file order; the base
state is the pose of the base *inertial* frame (COM); `link_com` is what `getLinkState()[0]`
returns.  Generalised velocity nu = [omega_world(3), v_basecom_world(3), qdot(n)], its time
derivative is the classical one.
"""
import xml.etree.ElementTree as ET
import numpy as np

from . import xform as xf


def _vec(s, n=3, default=None):
    if s is None:
        return np.array(default if default is not None else [0.0] * n, dtype=np.float64)
    return np.array([float(v) for v in s.split()], dtype=np.float64)


def skew(v):
    v = np.asarray(v)
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]
    S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2]
    S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]
    S[..., 2, 1] = v[..., 0]
    return S


class UrdfTree(object):
    def __init__(self, path):
        root = ET.parse(path).getroot()
        links = {}
        for l in root.findall("link"):
            inertial = l.find("inertial")
            if inertial is not None:
                o = inertial.find("origin")
                com = _vec(o.get("xyz") if o is not None else None)
                rpy = _vec(o.get("rpy") if o is not None else None)
                mass = float(inertial.find("mass").get("value"))
                it = inertial.find("inertia")
                g = lambda k: float(it.get(k, 0.0))
                I = np.array([[g("ixx"), g("ixy"), g("ixz")],
                              [g("ixy"), g("iyy"), g("iyz")],
                              [g("ixz"), g("iyz"), g("izz")]])
            else:
                com, rpy, mass, I = np.zeros(3), np.zeros(3), 0.0, np.zeros((3, 3))
            cols = []
            for c in l.findall("collision"):
                o = c.find("origin")
                cxyz = _vec(o.get("xyz") if o is not None else None)
                crpy = _vec(o.get("rpy") if o is not None else None)
                geom = c.find("geometry")
                if geom.find("sphere") is not None:
                    cols.append(("sphere", [float(geom.find("sphere").get("radius"))], cxyz, crpy))
                elif geom.find("box") is not None:
                    cols.append(("box", _vec(geom.find("box").get("size")).tolist(), cxyz, crpy))
                elif geom.find("cylinder") is not None or geom.find("capsule") is not None:
                    e = geom.find("cylinder") if geom.find("cylinder") is not None else geom.find("capsule")
                    kind = "cylinder" if geom.find("cylinder") is not None else "capsule"
                    cols.append((kind, [float(e.get("radius")), float(e.get("length"))], cxyz, crpy))
                else:
                    cols.append(("mesh", [], cxyz, crpy))
            fr = 0.5  # Bullet's default lateral friction when the URDF gives none
            ct = l.find("contact")
            if ct is not None and ct.find("lateral_friction") is not None:
                fr = float(ct.find("lateral_friction").get("value"))
            links[l.get("name")] = dict(com=com, R_in=xf.rpy_to_matrix(rpy), mass=mass, I=I, cols=cols, mu=fr)
        joints = []
        child_names = set()
        for j in root.findall("joint"):
            o = j.find("origin")
            ax = j.find("axis")
            lim = j.find("limit")
            dyn = j.find("dynamics")
            jt = j.get("type")
            joints.append(dict(
                name=j.get("name"), type=jt, parent=j.find("parent").get("link"),
                child=j.find("child").get("link"),
                p=_vec(o.get("xyz") if o is not None else None),
                R=xf.rpy_to_matrix(_vec(o.get("rpy") if o is not None else None)),
                axis=_vec(ax.get("xyz") if ax is not None else None, default=[1, 0, 0]),
                lower=float(lim.get("lower", 0.0)) if (lim is not None and jt == "revolute") else -np.inf,
                upper=float(lim.get("upper", 0.0)) if (lim is not None and jt == "revolute") else np.inf,
                damping=float(dyn.get("damping", 0.0)) if dyn is not None else 0.0))
            child_names.add(j.find("child").get("link"))
        roots = [n for n in links if n not in child_names]
        assert len(roots) == 1, roots
        self.base_name = roots[0]
        # depth-first, children in file order (PyBullet link indexing)
        order = []

        def visit(parent_name, parent_idx):
            for j in joints:
                if j["parent"] == parent_name:
                    idx = len(order)
                    order.append((j, parent_idx))
                    visit(j["child"], idx)
        visit(self.base_name, -1)
        self.n_links = len(order)
        self.joint_names = [j["name"] for j, _ in order]
        self.link_names = [j["child"] for j, _ in order]
        self.parent = np.array([p for _, p in order], dtype=np.int64)
        self.jtype = [j["type"] for j, _ in order]
        for t in self.jtype:
            if t not in ("revolute", "continuous", "fixed"):
                raise ValueError("Only support 1D and 0D joints at the moment.")  # imitation_task.py:871-879
        self.jp = np.stack([j["p"] for j, _ in order])
        self.jR = np.stack([j["R"] for j, _ in order])
        self.axis = np.stack([j["axis"] / np.linalg.norm(j["axis"]) for j, _ in order])
        self.lower = np.array([j["lower"] for j, _ in order])
        self.upper = np.array([j["upper"] for j, _ in order])
        self.damping = np.array([j["damping"] for j, _ in order])
        self.movable = np.array([t != "fixed" for t in self.jtype])
        self.qidx = -np.ones(self.n_links, dtype=np.int64)
        self.qidx[self.movable] = np.arange(int(self.movable.sum()))
        self.n_dof = int(self.movable.sum())
        bodies = [links[self.base_name]] + [links[n] for n in self.link_names]
        self.mass = np.array([b["mass"] for b in bodies])
        self.com = np.stack([b["com"] for b in bodies])
        self.R_in = np.stack([b["R_in"] for b in bodies])
        # inertia about the COM expressed in the LINK frame
        self.I_link = np.stack([b["R_in"] @ b["I"] @ b["R_in"].T for b in bodies])
        self.cols = [b["cols"] for b in bodies]
        self.mu = np.array([b["mu"] for b in bodies])
        # ancestors (movable joints on the path base -> link i, inclusive)
        self.anc = []
        for i in range(self.n_links):
            a, k = [], i
            while k >= 0:
                if self.movable[k]:
                    a.append(k)
                k = self.parent[k]
            self.anc.append(a[::-1])

    # -------------------------------------------------------------------------------------
    def fk(self, base_pos, base_quat, q):
        """Returns dict of world quantities, all batched (N, ...).

        R[b] link-frame rotation, x[b] link-frame origin, c[b] COM; b = 0 is the base, b = i+1 link i.
        """
        base_pos = np.atleast_2d(np.asarray(base_pos, dtype=np.float64))
        base_quat = np.atleast_2d(np.asarray(base_quat, dtype=np.float64))
        q = np.atleast_2d(np.asarray(q, dtype=np.float64))
        N = base_pos.shape[0]
        nb = self.n_links + 1
        R = np.zeros((N, nb, 3, 3))
        x = np.zeros((N, nb, 3))
        c = np.zeros((N, nb, 3))
        Rcom = xf.quat_to_matrix(base_quat)             # base inertial frame
        R[:, 0] = Rcom @ self.R_in[0].T                 # base link frame
        c[:, 0] = base_pos
        x[:, 0] = base_pos - np.einsum("nij,j->ni", R[:, 0], self.com[0])
        for i in range(self.n_links):
            p = self.parent[i] + 1
            Rj = R[:, p] @ self.jR[i]
            if self.movable[i]:
                Rj = Rj @ xf.axis_angle_matrix(self.axis[i], q[:, self.qidx[i]])
            R[:, i + 1] = Rj
            x[:, i + 1] = x[:, p] + np.einsum("nij,j->ni", R[:, p], self.jp[i])
            c[:, i + 1] = x[:, i + 1] + np.einsum("nij,j->ni", Rj, self.com[i + 1])
        return dict(R=R, x=x, c=c)

    def link_com(self, base_pos, base_quat, q):
        """World position of every link's COM frame, PyBullet getLinkState(...)[0]; (N, n_links, 3)."""
        return self.fk(base_pos, base_quat, q)["c"][:, 1:]

    # -------------------------------------------------------------------------------------
    def _world_axes(self, k):
        """World joint axes a[i] (N, n_links, 3): fixed in the child (and parent) link frame."""
        return np.einsum("nbij,bj->nbi", k["R"][:, 1:], self.axis)

    def point_jacobian(self, k, body, P):
        """Linear-velocity Jacobian (N,3,6+n) of the body-fixed point currently at world P (N,3)."""
        N = P.shape[0]
        J = np.zeros((N, 3, 6 + self.n_dof))
        J[:, :, 0:3] = -skew(P - k["c"][:, 0])
        J[:, :, 3:6] = np.eye(3)
        if body > 0:
            a = self._world_axes(k)
            for j in self.anc[body - 1]:
                J[:, :, 6 + self.qidx[j]] = np.cross(a[:, j], P - k["x"][:, j + 1])
        return J

    def body_jacobians(self, k):
        """J[b] (N, nb, 6, 6+n): [angular; linear COM velocity] of body b."""
        N = k["c"].shape[0]
        nb = self.n_links + 1
        J = np.zeros((N, nb, 6, 6 + self.n_dof))
        a = self._world_axes(k)
        for b in range(nb):
            J[:, b, 0:3, 0:3] = np.eye(3)
            J[:, b, 3:6, 0:3] = -skew(k["c"][:, b] - k["c"][:, 0])
            J[:, b, 3:6, 3:6] = np.eye(3)
            if b > 0:
                for j in self.anc[b - 1]:
                    J[:, b, 0:3, 6 + self.qidx[j]] = a[:, j]
                    J[:, b, 3:6, 6 + self.qidx[j]] = np.cross(a[:, j], k["c"][:, b] - k["x"][:, j + 1])
        return J

    def world_inertias(self, k, mass=None, inertia_scale=None):
        I = np.einsum("nbij,bjk,nblk->nbil", k["R"], self.I_link, k["R"])
        if inertia_scale is not None:
            I = I * inertia_scale[:, :, None, None]
        return I

    def mass_matrix(self, base_pos, base_quat, q, mass=None, inertia_scale=None):
        """M (N, 6+n, 6+n): kinetic energy T = 0.5 nu^T M nu."""
        k = self.fk(base_pos, base_quat, q)
        J = self.body_jacobians(k)
        Iw = self.world_inertias(k, inertia_scale=inertia_scale)
        m = self.mass if mass is None else mass
        m = np.broadcast_to(m, Iw.shape[:2])
        Ja, Jl = J[:, :, 0:3], J[:, :, 3:6]
        M = np.einsum("nbia,nbij,nbjc->nac", Ja, Iw, Ja) + np.einsum("nb,nbia,nbic->nac", m, Jl, Jl)
        return M

    def bias(self, base_pos, base_quat, q, nu, gravity=(0.0, 0.0, -10.0), mass=None, inertia_scale=None, k=None):
        """C (N, 6+n) such that M nudot + C = generalised applied forces (incl. gravity in C)."""
        if k is None:
            k = self.fk(base_pos, base_quat, q)
        nu = np.atleast_2d(np.asarray(nu, dtype=np.float64))
        N = nu.shape[0]
        nb = self.n_links + 1
        g = np.asarray(gravity, dtype=np.float64)
        a_ax = self._world_axes(k)
        w = np.zeros((N, nb, 3))
        al = np.zeros((N, nb, 3))
        ax = np.zeros((N, nb, 3))   # velocity-product acceleration of the link origin
        ac = np.zeros((N, nb, 3))   # ... of the COM
        w[:, 0] = nu[:, 0:3]
        r0 = k["x"][:, 0] - k["c"][:, 0]
        ax[:, 0] = np.cross(w[:, 0], np.cross(w[:, 0], r0))
        for i in range(self.n_links):
            p = self.parent[i] + 1
            r = k["x"][:, i + 1] - k["x"][:, p]
            ax[:, i + 1] = ax[:, p] + np.cross(al[:, p], r) + np.cross(w[:, p], np.cross(w[:, p], r))
            if self.movable[i]:
                wj = a_ax[:, i] * nu[:, 6 + self.qidx[i]][:, None]
                w[:, i + 1] = w[:, p] + wj
                al[:, i + 1] = al[:, p] + np.cross(w[:, p], wj)
            else:
                w[:, i + 1] = w[:, p]
                al[:, i + 1] = al[:, p]
            rc = k["c"][:, i + 1] - k["x"][:, i + 1]
            ac[:, i + 1] = ax[:, i + 1] + np.cross(al[:, i + 1], rc) + np.cross(w[:, i + 1], np.cross(w[:, i + 1], rc))
        Iw = self.world_inertias(k, inertia_scale=inertia_scale)
        m = np.broadcast_to(self.mass if mass is None else mass, Iw.shape[:2])
        Iw_w = np.einsum("nbij,nbj->nbi", Iw, w)
        n_b = np.einsum("nbij,nbj->nbi", Iw, al) + np.cross(w, Iw_w)
        f_b = m[..., None] * (ac - g)
        J = self.body_jacobians(k)
        C = np.einsum("nbia,nbi->na", J[:, :, 0:3], n_b) + np.einsum("nbia,nbi->na", J[:, :, 3:6], f_b)
        return C

    def inverse_dynamics(self, base_pos, base_quat, q, nu, nudot, gravity=(0.0, 0.0, -10.0)):
        M = self.mass_matrix(base_pos, base_quat, q)
        C = self.bias(base_pos, base_quat, q, nu, gravity)
        return np.einsum("nab,nb->na", M, np.atleast_2d(nudot)) + C

    def kinetic_energy(self, base_pos, base_quat, q, nu):
        M = self.mass_matrix(base_pos, base_quat, q)
        nu = np.atleast_2d(nu)
        return 0.5 * np.einsum("na,nab,nb->n", nu, M, nu)

    def potential_energy(self, base_pos, base_quat, q, gravity=(0.0, 0.0, -10.0)):
        k = self.fk(base_pos, base_quat, q)
        return -np.einsum("b,nbi,i->n", self.mass, k["c"], np.asarray(gravity))

    # -------------------------------------------------------------------------------------
    def collision_points(self):
        """Sphere-set approximation of the collision primitives.

        Returns list of (body, local point (link frame), radius).  sphere -> itself; box -> its 8
        corners (radius 0: exact for box-vs-plane); cylinder/capsule -> the two end-cap centres
        with the shape radius; mesh -> ignored.  This is the geometric contact criterion of
        SURVEY.md section 7 ("any collision primitive within the contact threshold of z=0").
        """
        pts = []
        for b, cols in enumerate(self.cols):
            for kind, prm, cxyz, crpy in cols:
                Rc = xf.rpy_to_matrix(crpy)
                if kind == "sphere":
                    pts.append((b, cxyz.copy(), prm[0]))
                elif kind == "box":
                    h = np.array(prm) / 2.0
                    for sx in (-1, 1):
                        for sy in (-1, 1):
                            for sz in (-1, 1):
                                pts.append((b, cxyz + Rc @ (h * np.array([sx, sy, sz])), 0.0))
                elif kind in ("cylinder", "capsule"):
                    for sz in (-1, 1):
                        pts.append((b, cxyz + Rc @ np.array([0, 0, sz * prm[1] / 2.0]), prm[0]))
        return pts
