"""Independent derivation of every compiled-model constant from the MJCF text (numpy + ElementTree).

TEST INFRASTRUCTURE ONLY.  The product compiles MJCF with csrc/mjcf_compile.cpp and the CPU oracle consumes
that same blob, so a wrong constant (an inertia formula, `invweight0`, the pair filter, solref mixing) would be
invisible to every CUDA-vs-oracle comparison.  This module re-derives the blob's fields a second time, from the
XML alone, sharing no code with the compiler: rotation matrices instead of quaternion algebra, inertias by the
parallel-axis construction, `invweight0` from a dense `J M^-1 J^T` built with numpy, the pair table by nested
body loops.  tests/test_model_constants.py asserts blob == this derivation to 1e-12 for every scene.

What is being restated (the reference's call site is `mj.MjModel.from_xml_path`, MuJoCo_Gym/mujoco_parent.py:126;
the rules are those of MuJoCo's documented compiler / `mj_setConst`, mujoco==2.3.3, for the MJCF subset of
SURVEY.md A.2):
  * local coordinates, `angle` degree|radian, `eulerseq`; `fromto` -> pos / minimal z-to-vector rotation / half length
  * inertiafromgeom: mass = density * volume, principal inertias of sphere / capsule / box about the geom frame
  * qpos0 = body pos / quat for free joints, `ref` for hinge / slide
  * weld ids, tree ids, dof parents; bounding radii
  * collision filter: contype / conaffinity, same weld body, weld parent-child unless one is the world
  * pair mixing: margin / gap max, friction max, condim max, solref / solimp by solmix weights
  * body_invweight0 = tr(J M^-1 J^T) / 3 (translation | rotation blocks) at the body's inertial frame, qpos0;
    dof_invweight0 = diag(M^-1), averaged over the 3 translations / 3 rotations of a free joint
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

PLANE, SPHERE, CAPSULE, BOX = 0, 2, 3, 6
FREE, SLIDE, HINGE = 0, 2, 3
SENSORS = {"touch": (0, 1, 1), "accelerometer": (1, 3, 0), "rangefinder": (7, 1, 1),
           "framexaxis": (28, 3, 2), "frameyaxis": (29, 3, 2), "framezaxis": (30, 3, 2)}


def _nums(s):
    return [float(x) for x in s.replace(",", " ").split()]


def _rot_axis(axis, ang):
    """Rodrigues rotation matrix"""
    a = np.asarray(axis, float)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def mat2quat(R):
    """rotation matrix -> (w, x, y, z) with w >= 0 (largest-component branch for stability)"""
    t = np.trace(R)
    cand = [t, R[0, 0] - R[1, 1] - R[2, 2], R[1, 1] - R[0, 0] - R[2, 2], R[2, 2] - R[0, 0] - R[1, 1]]
    k = int(np.argmax(cand))
    s = math.sqrt(max(0.0, 1.0 + cand[k])) * 2.0
    if k == 0:
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif k == 1:
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif k == 2:
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    q /= np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def quat2mat(q):
    w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _z_to(vec):
    """minimal rotation taking +z onto vec (mju_quatZ2Vec semantics: axis z x vec, x axis when degenerate)"""
    v = np.asarray(vec, float)
    v = v / np.linalg.norm(v)
    ax = np.cross([0, 0, 1.0], v)
    s = np.linalg.norm(ax)
    ang = math.atan2(s, v[2])
    if s < 1e-10:
        ax = np.array([1.0, 0, 0])
    return _rot_axis(ax, ang)


class RefModel:
    """fields named as in the packed blob (include/mjb_blob.h); every array is a flat numpy array"""

    def __init__(self, xml_text):
        root = ET.fromstring(xml_text)
        self.f = {}
        comp = root.find("compiler")
        self.deg = True
        self.eulerseq = "xyz"
        if comp is not None:
            self.deg = comp.get("angle", "degree") == "degree"
            self.eulerseq = comp.get("eulerseq", "xyz")
        opt = root.find("option")
        og = (lambda k, d: opt.get(k, d) if opt is not None else d)
        self.timestep = float(og("timestep", "0.002"))
        self.integrator = {"Euler": 0, "RK4": 1}[og("integrator", "Euler")]
        self.gravity = np.array(_nums(og("gravity", "0 0 -9.81")))
        dflt = root.find("default")
        self.dflt = {k: (dflt.find(k) if dflt is not None else None) for k in ("joint", "geom", "site", "motor")}
        self.bodies, self.joints, self.geoms, self.sites = [], [], [], []
        wb = root.find("worldbody")
        self.bodies.append({"name": "world", "parent": 0, "R": np.eye(3), "pos": np.zeros(3), "joints": [], "geoms": []})
        for el in wb:
            if el.tag == "geom":
                self._geom(el, 0)
            elif el.tag == "site":
                self._site(el, 0)
        for el in wb:
            if el.tag == "body":
                self._body(el, 0)
        self.sensors = []
        sn = root.find("sensor")
        if sn is not None:
            names = [s["name"] for s in self.sites]
            for el in sn:
                t, dim, dt = SENSORS[el.tag]
                site = el.get("objname") if dt == 2 else el.get("site")
                self.sensors.append({"type": t, "dim": dim, "dtype": dt, "site": names.index(site), "cutoff": float(el.get("cutoff", 0))})
        self.motors = []
        an = root.find("actuator")
        if an is not None:
            jn = [j["name"] for j in self.joints]
            for el in an:
                g = self._attr(el, "motor")
                cr = _nums(g("ctrlrange", "0 0"))
                cl = g("ctrllimited", None)
                lim = cl == "true" or (cl == "auto" and g("ctrlrange", None) is not None)
                self.motors.append({"joint": jn.index(el.get("joint")), "gear": _nums(g("gear", "1"))[0], "range": cr, "limited": int(lim)})
        self._derive()

    # -- attribute lookup with the single unnamed default class
    def _attr(self, el, kind):
        d = self.dflt.get(kind)

        def get(k, default):
            v = el.get(k)
            if v is None and d is not None:
                v = d.get(k)
            return default if v is None else v
        return get

    def _orient(self, el):
        """local rotation matrix from quat | euler | axisangle | xyaxes | zaxis (element attributes only)"""
        if el.get("quat"):
            return quat2mat(_nums(el.get("quat")))
        if el.get("euler"):
            R = np.eye(3)
            for ang, c in zip(_nums(el.get("euler")), self.eulerseq):
                a = math.radians(ang) if self.deg else ang
                E = _rot_axis(np.eye(3)["xyz".index(c.lower())], a)
                R = R @ E if c.islower() else E @ R   # intrinsic: about the rotated axes; extrinsic: about the fixed ones
            return R
        if el.get("axisangle"):
            v = _nums(el.get("axisangle"))
            return _rot_axis(v[:3], math.radians(v[3]) if self.deg else v[3])
        if el.get("xyaxes"):
            v = np.array(_nums(el.get("xyaxes")))
            x = v[:3] / np.linalg.norm(v[:3])
            y = v[3:] - x * (x @ v[3:])
            y /= np.linalg.norm(y)
            return np.stack([x, y, np.cross(x, y)], axis=1)
        if el.get("zaxis"):
            return _z_to(_nums(el.get("zaxis")))
        return np.eye(3)

    def _body(self, el, parent):
        bid = len(self.bodies)
        b = {"name": el.get("name", ""), "parent": parent, "pos": np.array(_nums(el.get("pos", "0 0 0"))), "R": self._orient(el),
             "joints": [], "geoms": []}
        self.bodies.append(b)
        for c in el:
            if c.tag in ("joint", "freejoint"):
                self._joint(c, bid)
            elif c.tag == "geom":
                self._geom(c, bid)
            elif c.tag == "site":
                self._site(c, bid)
        for c in el:
            if c.tag == "body":
                self._body(c, bid)

    def _joint(self, el, bid):
        free = el.tag == "freejoint" or el.get("type") == "free" or (el.get("type") is None and self.dflt["joint"] is not None and
                                                                   self.dflt["joint"].get("type") == "free")
        g = (lambda k, d: el.get(k, d)) if el.tag == "freejoint" else self._attr(el, "joint")
        typ = FREE if free else {"hinge": HINGE, "slide": SLIDE}[g("type", "hinge")]
        ax = np.array(_nums(g("axis", "0 0 1")))
        rng = _nums(g("range", "0 0"))
        ref = float(g("ref", "0"))
        if self.deg and typ == HINGE:
            rng = [math.radians(r) for r in rng]
            ref = math.radians(ref)
        lim = g("limited", None)
        limited = (lim == "true") or (lim == "auto" and g("range", None) is not None)
        j = {"name": el.get("name", ""), "type": typ, "body": bid, "pos": np.zeros(3) if typ == FREE else np.array(_nums(g("pos", "0 0 0"))),
             "axis": np.array([0, 0, 1.0]) if typ == FREE else ax / np.linalg.norm(ax), "range": rng, "limited": int(limited and typ != FREE),
             "margin": float(g("margin", "0")), "armature": float(g("armature", "0")), "damping": float(g("damping", "0")), "ref": ref,
             "solref": _nums(g("solreflimit", "0.02 1")), "solimp": (_nums(g("solimplimit", "0.9 0.95 0.001 0.5 2")) + [0.001, 0.5, 2])[:5]}
        self.bodies[bid]["joints"].append(len(self.joints))
        self.joints.append(j)

    def _geom(self, el, bid):
        g = self._attr(el, "geom")
        typ = {"plane": PLANE, "sphere": SPHERE, "capsule": CAPSULE, "box": BOX}[g("type", "sphere")]
        size = (_nums(g("size", "0 0 0")) + [0, 0, 0])[:3]
        if el.get("fromto"):
            ft = np.array(_nums(el.get("fromto")))
            d = ft[3:] - ft[:3]
            pos, R = 0.5 * (ft[:3] + ft[3:]), _z_to(d)
            size[1 if typ == CAPSULE else 2] = 0.5 * np.linalg.norm(d)
        else:
            pos, R = np.array(_nums(g("pos", "0 0 0"))), self._orient(el)
        pad = lambda v, d: (v + d[len(v):])[:len(d)]
        rec = {"name": el.get("name", ""), "type": typ, "body": bid, "size": np.array(size, float), "pos": pos, "R": R,
               "contype": int(float(g("contype", "1"))), "conaffinity": int(float(g("conaffinity", "1"))), "condim": int(float(g("condim", "3"))),
               "friction": pad(_nums(g("friction", "1 0.005 0.0001")), [1, 0.005, 0.0001]), "margin": float(g("margin", "0")),
               "gap": float(g("gap", "0")), "solmix": float(g("solmix", "1")), "solref": _nums(g("solref", "0.02 1")),
               "solimp": pad(_nums(g("solimp", "0.9 0.95 0.001 0.5 2")), [0.9, 0.95, 0.001, 0.5, 2]),
               "rgba": _nums(g("rgba", "0.5 0.5 0.5 1")), "density": float(g("density", "1000")),
               "mass": float(el.get("mass")) if el.get("mass") is not None else None}
        self.bodies[bid]["geoms"].append(len(self.geoms))
        self.geoms.append(rec)

    def _site(self, el, bid):
        g = self._attr(el, "site")
        size = _nums(g("size", "0.005 0.005 0.005"))
        size = (size + [0.005, 0.005, 0.005][len(size):])[:3]
        self.sites.append({"name": el.get("name", ""), "body": bid, "pos": np.array(_nums(g("pos", "0 0 0"))), "R": self._orient(el),
                           "type": {"sphere": SPHERE, "capsule": CAPSULE, "box": BOX}[g("type", "sphere")], "size": size})

    # -- inertia of one geom about its own centre, in its own axes
    @staticmethod
    def geom_inertia(typ, size, density, mass_attr):
        if typ == SPHERE:
            r = size[0]
            m = density * 4.0 / 3.0 * math.pi * r ** 3 if mass_attr is None else mass_attr
            return m, np.full(3, 0.4 * m * r * r)
        if typ == BOX:
            a, b, c = 2 * size[0], 2 * size[1], 2 * size[2]    # full edge lengths
            m = density * a * b * c if mass_attr is None else mass_attr
            return m, np.array([m * (b * b + c * c) / 12, m * (a * a + c * c) / 12, m * (a * a + b * b) / 12])
        if typ == CAPSULE:
            r, L = size[0], 2 * size[1]
            vc, vs = math.pi * r * r * L, 4.0 / 3.0 * math.pi * r ** 3     # cylinder, the two hemispheres together
            m = density * (vc + vs) if mass_attr is None else mass_attr
            mc, mh = m * vc / (vc + vs), 0.5 * m * vs / (vc + vs)           # cylinder mass, ONE hemisphere's mass
            izz = 0.5 * mc * r * r + 2 * (0.4 * mh * r * r)
            # transverse: cylinder + 2 hemispheres.  A hemisphere about a diameter of its flat face has 2/5 m r^2; its
            # centre of mass is 3r/8 above that face: shift to the com, then out to the capsule centre (L/2 + 3r/8)
            ih_com = 0.4 * mh * r * r - mh * (3 * r / 8) ** 2
            ixx = mc * (3 * r * r + L * L) / 12 + 2 * (ih_com + mh * (L / 2 + 3 * r / 8) ** 2)
            return m, np.array([ixx, ixx, izz])
        return 0.0, np.zeros(3)

    def _derive(self):
        B, J, G = self.bodies, self.joints, self.geoms
        nb = len(B)
        f = self.f
        # qpos / dof layout
        nq = nv = 0
        for j in J:
            j["qadr"], j["dadr"] = nq, nv
            nq += 7 if j["type"] == FREE else 1
            nv += 6 if j["type"] == FREE else 1
        self.nq, self.nv = nq, nv
        qpos0 = np.zeros(nq)
        dof_body, dof_jnt, dof_parent = np.zeros(nv, int), np.zeros(nv, int), np.zeros(nv, int)
        dof_arm, dof_damp = np.zeros(nv), np.zeros(nv)
        last = [-1] * nb
        weld, rootb, depth = [0] * nb, [0] * nb, [0] * nb
        for b in range(1, nb):
            p = B[b]["parent"]
            rootb[b] = b if p == 0 else rootb[p]
            weld[b] = b if B[b]["joints"] else weld[p]
            depth[b] = depth[p] + 1
            cur = last[p]
            for ji in B[b]["joints"]:
                j = J[ji]
                nd = 6 if j["type"] == FREE else 1
                for i in range(nd):
                    d = j["dadr"] + i
                    dof_body[d], dof_jnt[d], dof_parent[d] = b, ji, cur
                    dof_arm[d], dof_damp[d] = j["armature"], j["damping"]
                    cur = d
                if j["type"] == FREE:
                    qpos0[j["qadr"]:j["qadr"] + 3] = B[b]["pos"]
                    qpos0[j["qadr"] + 3:j["qadr"] + 7] = mat2quat(B[b]["R"])
                else:
                    qpos0[j["qadr"]] = j["ref"]
            last[b] = cur
        has_dof = [bool(B[b]["joints"]) for b in range(nb)]
        for b in range(nb - 1, 0, -1):
            if has_dof[b]:
                has_dof[B[b]["parent"]] = True
        tree_of_root, ntree = {}, 0
        for b in range(1, nb):
            if B[b]["parent"] == 0 and has_dof[b]:
                tree_of_root[b] = ntree
                ntree += 1
        treeid = [-1] + [tree_of_root.get(rootb[b], -1) for b in range(1, nb)]
        # inertial properties
        mass, inertia, ipos, iR = np.zeros(nb), np.zeros((nb, 3)), np.zeros((nb, 3)), [np.eye(3) for _ in range(nb)]
        for b in range(1, nb):
            gs = B[b]["geoms"]
            if not gs:
                continue
            mi = [self.geom_inertia(G[g]["type"], G[g]["size"], G[g]["density"], G[g]["mass"]) for g in gs]
            mt = sum(m for m, _ in mi)
            if mt <= 0:
                continue
            com = sum(G[g]["pos"] * m for g, (m, _) in zip(gs, mi)) / mt
            mass[b], ipos[b] = mt, com
            if len(gs) == 1:
                inertia[b], iR[b] = mi[0][1], G[gs[0]]["R"]
            else:
                T = np.zeros((3, 3))
                for g, (m, I) in zip(gs, mi):
                    R, d = G[g]["R"], G[g]["pos"] - com
                    T += R @ np.diag(I) @ R.T + m * ((d @ d) * np.eye(3) - np.outer(d, d))
                w, V = np.linalg.eigh(T)
                order = np.argsort(-w)
                inertia[b], V = w[order], V[:, order]
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                iR[b] = V
        subtree = mass.copy()
        for b in range(nb - 1, 0, -1):
            subtree[B[b]["parent"]] += subtree[b]
        # kinematics at qpos0 (hinge / slide at their reference: the body sits at its MJCF pose)
        xR, xp = [np.eye(3)] * nb, [np.zeros(3)] * nb
        for b in range(1, nb):
            p = B[b]["parent"]
            xR[b], xp[b] = xR[p] @ B[b]["R"], xp[p] + xR[p] @ B[b]["pos"]
        # dense Jacobians: column d = velocity of a point / angular velocity per unit qvel[d]
        def jac(b, point):
            jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
            bb = b
            while bb > 0:
                for ji in B[bb]["joints"]:
                    j = J[ji]
                    d = j["dadr"]
                    if j["type"] == FREE:
                        jp[:, d:d + 3] = np.eye(3)                       # world-frame translation
                        for i in range(3):                                # body-frame rotation axes
                            w = xR[bb][:, i]
                            jr[:, d + 3 + i], jp[:, d + 3 + i] = w, np.cross(w, point - xp[bb])
                    elif j["type"] == HINGE:
                        w, anchor = xR[bb] @ j["axis"], xp[bb] + xR[bb] @ j["pos"]
                        jr[:, d], jp[:, d] = w, np.cross(w, point - anchor)
                    else:
                        jp[:, d] = xR[bb] @ j["axis"]
                bb = B[bb]["parent"]
            return jp, jr
        M = np.diag(dof_arm).astype(float)
        for b in range(1, nb):
            if mass[b] <= 0:
                continue
            c = xp[b] + xR[b] @ ipos[b]
            jp, jr = jac(b, c)
            Rw = xR[b] @ iR[b]
            M += mass[b] * jp.T @ jp + jr.T @ (Rw @ np.diag(inertia[b]) @ Rw.T) @ jr
        self.M0 = M
        body_iw, dof_iw = np.zeros((nb, 2)), np.zeros(nv)
        if nv:
            Minv = np.linalg.inv(M)
            for b in range(1, nb):
                if mass[b] < 1e-15 or treeid[b] < 0:
                    continue
                jp, jr = jac(b, xp[b] + xR[b] @ ipos[b])
                body_iw[b] = [max(1e-15, np.trace(jp @ Minv @ jp.T) / 3), max(1e-15, np.trace(jr @ Minv @ jr.T) / 3)]
            dg = np.diag(Minv)
            for j in J:
                d = j["dadr"]
                if j["type"] == FREE:
                    dof_iw[d:d + 3], dof_iw[d + 3:d + 6] = dg[d:d + 3].mean(), dg[d + 3:d + 6].mean()
                else:
                    dof_iw[d] = dg[d]
        # collision pair table
        pairs = []
        for b1 in range(nb):
            for b2 in range(b1 + 1, nb):
                w1, w2 = weld[b1], weld[b2]
                if w1 == w2:
                    continue
                if w1 and w2 and (weld[B[w1]["parent"]] == w2 or weld[B[w2]["parent"]] == w1):
                    continue
                for g1 in B[b1]["geoms"]:
                    for g2 in B[b2]["geoms"]:
                        a, c = G[g1], G[g2]
                        if not ((a["contype"] & c["conaffinity"]) or (c["contype"] & a["conaffinity"])):
                            continue
                        if a["type"] == PLANE and c["type"] == PLANE:
                            continue
                        pairs.append((g1, g2) if a["type"] <= c["type"] else (g2, g1))
        pm, pim, pf, psr, psi, pcd = [], [], [], [], [], []
        for g1, g2 in pairs:
            a, c = G[g1], G[g2]
            pcd.append(max(a["condim"], c["condim"]))
            m_, gp = max(a["margin"], c["margin"]), max(a["gap"], c["gap"])
            pm.append(m_)
            pim.append(m_ - gp)
            pf.append(np.maximum(a["friction"], c["friction"]))
            s1, s2 = a["solmix"], c["solmix"]
            tiny = 1e-15
            mix = s1 / (s1 + s2) if (s1 >= tiny and s2 >= tiny) else (0.5 if (s1 < tiny and s2 < tiny) else (0.0 if s1 < tiny else 1.0))
            r1, r2 = np.array(a["solref"]), np.array(c["solref"])
            psr.append(mix * r1 + (1 - mix) * r2 if (r1[0] > 0 and r2[0] > 0) else np.minimum(r1, r2))
            psi.append(mix * np.array(a["solimp"]) + (1 - mix) * np.array(c["solimp"]))
        rb = {PLANE: lambda s: 0.0, SPHERE: lambda s: s[0], CAPSULE: lambda s: s[0] + s[1], BOX: lambda s: float(np.linalg.norm(s))}
        I = lambda x: np.asarray(x, dtype=np.int64).reshape(-1)
        F = lambda x: np.asarray(x, dtype=np.float64).reshape(-1)
        f.update({
            "nq": I([nq]), "nv": I([nv]), "nu": I([len(self.motors)]), "nbody": I([nb]), "njnt": I([len(J)]), "ngeom": I([len(G)]),
            "nsite": I([len(self.sites)]), "nsensor": I([len(self.sensors)]), "nsensordata": I([sum(s["dim"] for s in self.sensors)]),
            "npair": I([len(pairs)]), "opt_integrator": I([self.integrator]), "ntree": I([ntree]), "maxdepth": I([max(depth)]),
            "opt_timestep": F([self.timestep]), "opt_gravity": F(self.gravity),
            "body_parentid": I([b["parent"] for b in B]), "body_rootid": I(rootb), "body_weldid": I(weld),
            "body_jntnum": I([len(b["joints"]) for b in B]), "body_jntadr": I([b["joints"][0] if b["joints"] else -1 for b in B]),
            "body_dofnum": I([sum(6 if J[j]["type"] == FREE else 1 for j in b["joints"]) for b in B]),
            "body_dofadr": I([J[b["joints"][0]]["dadr"] if b["joints"] else -1 for b in B]),
            "body_geomnum": I([len(b["geoms"]) for b in B]), "body_geomadr": I([b["geoms"][0] if b["geoms"] else -1 for b in B]),
            "body_depth": I(depth), "body_treeid": I(treeid),
            "body_pos": F([b["pos"] for b in B]), "body_quat": F([mat2quat(b["R"]) for b in B]), "body_ipos": F(ipos),
            "body_iquat": F([mat2quat(r) for r in iR]), "body_mass": F(mass), "body_inertia": F(inertia), "body_subtreemass": F(subtree),
            "body_invweight0": F(body_iw),
            "jnt_type": I([j["type"] for j in J]), "jnt_bodyid": I([j["body"] for j in J]), "jnt_qposadr": I([j["qadr"] for j in J]),
            "jnt_dofadr": I([j["dadr"] for j in J]), "jnt_limited": I([j["limited"] for j in J]),
            "jnt_pos": F([j["pos"] for j in J]), "jnt_axis": F([j["axis"] for j in J]), "jnt_range": F([j["range"] for j in J]),
            "jnt_margin": F([j["margin"] for j in J]), "jnt_solref": F([j["solref"] for j in J]), "jnt_solimp": F([j["solimp"] for j in J]),
            "dof_bodyid": I(dof_body), "dof_jntid": I(dof_jnt), "dof_parentid": I(dof_parent), "dof_armature": F(dof_arm),
            "dof_damping": F(dof_damp), "dof_invweight0": F(dof_iw),
            "geom_type": I([g["type"] for g in G]), "geom_bodyid": I([g["body"] for g in G]), "geom_contype": I([g["contype"] for g in G]),
            "geom_conaffinity": I([g["conaffinity"] for g in G]), "geom_condim": I([g["condim"] for g in G]),
            "geom_size": F([g["size"] for g in G]), "geom_pos": F([g["pos"] for g in G]), "geom_quat": F([mat2quat(g["R"]) for g in G]),
            "geom_friction": F([g["friction"] for g in G]), "geom_margin": F([g["margin"] for g in G]), "geom_gap": F([g["gap"] for g in G]),
            "geom_solmix": F([g["solmix"] for g in G]), "geom_solref": F([g["solref"] for g in G]), "geom_solimp": F([g["solimp"] for g in G]),
            "geom_rbound": F([rb[g["type"]](g["size"]) for g in G]), "geom_rgba": F([g["rgba"] for g in G]),
            "site_bodyid": I([s["body"] for s in self.sites]), "site_type": I([s["type"] for s in self.sites]),
            "site_pos": F([s["pos"] for s in self.sites]), "site_quat": F([mat2quat(s["R"]) for s in self.sites]),
            "site_size": F([s["size"] for s in self.sites]),
            "sensor_type": I([s["type"] for s in self.sensors]), "sensor_objid": I([s["site"] for s in self.sensors]),
            "sensor_adr": I(np.cumsum([0] + [s["dim"] for s in self.sensors])[:-1]), "sensor_dim": I([s["dim"] for s in self.sensors]),
            "sensor_datatype": I([s["dtype"] for s in self.sensors]), "sensor_cutoff": F([s["cutoff"] for s in self.sensors]),
            "actuator_trnid": I([m["joint"] for m in self.motors]), "actuator_ctrllimited": I([m["limited"] for m in self.motors]),
            "actuator_gear": F([m["gear"] for m in self.motors]), "actuator_ctrlrange": F([m["range"] for m in self.motors]),
            "pair_geom1": I([p[0] for p in pairs]), "pair_geom2": I([p[1] for p in pairs]), "pair_condim": I(pcd),
            "pair_margin": F(pm), "pair_includemargin": F(pim), "pair_friction": F(pf), "pair_solref": F(psr), "pair_solimp": F(psi),
            "qpos0": F(qpos0),
        })

    QUAT_FIELDS = ("body_quat", "body_iquat", "geom_quat", "site_quat")
