"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64) of the MANO hand layer the reference calls at
network/gen_net.py:116-118 (`self.rh_mano(betas=..., global_orient=..., hand_pose=..., transl=...).vertices`), created at
gen_diverse_grasp_obman.py:355-360 with `mano.load(model_path='./models/mano/MANO_RIGHT.pkl', model_type='mano',
use_pca=True, num_pca_comps=45, batch_size=1, flat_hand_mean=True)`.

The arithmetic lives in a third-party package that is NOT vendored in the reference (`mano`, otaheri/MANO, version unpinned: no
requirements file; it wraps the linear-blend-skinning of `smplx/lbs.py`).  This file restates that published algorithm:

    hand_pose  = pca_coeffs @ hands_components[:ncomps]                       (use_pca)
    full_pose  = [global_orient | hand_pose] + pose_mean                      (pose_mean = 0 with flat_hand_mean=True)
    v_shaped   = v_template + shapedirs . betas
    J          = J_regressor @ v_shaped
    R_j        = rodrigues(full_pose_j)            angle = |r + 1e-8|, R = I + sin K + (1 - cos) K K
    v_posed    = v_shaped + posedirs . vec(R_1..15 - I)
    G_j        = G_parent(j) @ [R_j | J_j - J_parent(j)],   A_j = [G_j.R | G_j.t - G_j.R J_j]
    vertices   = (sum_j w_vj A_j) [v_posed; 1] + transl

PARITY: unpinned against the `mano` package itself (absent here, cannot be executed); pinned against the reference's own asset
where that is possible — `models/mano/MANO_RIGHT.pkl` is readable with a stub for the (absent) `chumpy` classes it pickles, and
tests/test_mano.py checks the invariants the asset carries (zero pose and shape reproduce `v_template`; `J_regressor @
v_template` equals the pickled rest joints `J`; the kinematic tree is the 5 x 3 finger chain).  Only tests/, smoke() and
bench.py may import this module.
"""
import pickle
import sys
import types

import numpy as np

N_VERTS, N_JOINTS, N_BETAS, N_POSE = 778, 16, 10, 45


def load_pkl(path):
    """MANO_{LEFT,RIGHT}.pkl -> dict of float64 numpy arrays (the pickle holds chumpy objects: read through a stub)."""
    class Ch:
        def __setstate__(self, st):
            self.__dict__.update(st if isinstance(st, dict) else {"state": st})

    saved = {k: sys.modules.get(k) for k in ("chumpy", "chumpy.ch", "chumpy.reordering", "chumpy.ch_ops", "chumpy.logic", "chumpy.utils")}
    try:
        for name in saved:
            m = types.ModuleType(name)
            m.Ch = Ch
            m.__getattr__ = lambda n, _C=Ch: type(n, (_C,), {})
            sys.modules[name] = m
        with open(path, "rb") as f:
            d = pickle.load(f, encoding="latin1")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    def arr(v):
        if isinstance(v, np.ndarray):
            return np.asarray(v, dtype=np.float64)
        if hasattr(v, "toarray"):
            return np.asarray(v.toarray(), dtype=np.float64)
        st = v.__dict__
        if "x" in st:                                   # chumpy.Ch: the value
            return np.asarray(st["x"], dtype=np.float64)
        if "a" in st and "idxs" in st:                  # chumpy.reordering.Select: a.ravel()[idxs]
            return arr(st["a"]).ravel()[np.asarray(st["idxs"])]
        raise TypeError("cannot read %r" % (v,))

    return {
        "v_template": arr(d["v_template"]).reshape(N_VERTS, 3),
        "shapedirs": arr(d["shapedirs"]).reshape(N_VERTS, 3, N_BETAS),
        "posedirs": arr(d["posedirs"]).reshape(N_VERTS, 3, 9 * (N_JOINTS - 1)),
        "J_regressor": arr(d["J_regressor"]).reshape(N_JOINTS, N_VERTS),
        "weights": arr(d["weights"]).reshape(N_VERTS, N_JOINTS),
        "hands_components": arr(d["hands_components"]).reshape(N_POSE, N_POSE),
        "hands_mean": arr(d["hands_mean"]).reshape(N_POSE),
        "parents": np.asarray(d["kintree_table"])[0].astype(np.int64),     # parents[0] is the root's marker (2^32 - 1)
        "J": arr(d["J"]).reshape(N_JOINTS, 3),
        "faces": np.asarray(d["f"]).astype(np.int64),
    }


def synthetic_model(seed=0):
    """A structurally valid random model (same shapes, MANO's kinematic tree, convex skinning weights) for tests that must run
    without the asset (the GPU box has no /root/reference)."""
    rs = np.random.RandomState(seed)
    w = rs.rand(N_VERTS, N_JOINTS) ** 4
    w /= w.sum(1, keepdims=True)
    jr = rs.rand(N_JOINTS, N_VERTS) ** 8
    jr /= jr.sum(1, keepdims=True)
    return {
        "v_template": 0.1 * rs.randn(N_VERTS, 3),
        "shapedirs": 0.01 * rs.randn(N_VERTS, 3, N_BETAS),
        "posedirs": 0.002 * rs.randn(N_VERTS, 3, 9 * (N_JOINTS - 1)),
        "J_regressor": jr,
        "weights": w,
        "hands_components": rs.randn(N_POSE, N_POSE) / np.sqrt(N_POSE),
        "hands_mean": 0.3 * rs.randn(N_POSE),
        "parents": np.array([-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14], dtype=np.int64),
        "faces": rs.randint(0, N_VERTS, size=(1538, 3)).astype(np.int64),
    }


def rodrigues(rot_vecs):
    """smplx.lbs.batch_rodrigues: [n,3] axis-angle -> [n,3,3]; the epsilon is added to every component before the norm."""
    r = np.asarray(rot_vecs, dtype=np.float64)
    angle = np.linalg.norm(r + 1e-8, axis=1, keepdims=True)
    d = r / angle
    c, s = np.cos(angle)[:, :, None], np.sin(angle)[:, :, None]
    K = np.zeros((r.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -d[:, 2], d[:, 1]
    K[:, 1, 0], K[:, 1, 2] = d[:, 2], -d[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -d[:, 1], d[:, 0]
    return np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)


def mano_forward(model, betas, global_orient=None, hand_pose=None, transl=None, use_pca=True, num_pca_comps=45, flat_hand_mean=True):
    """-> (vertices [B,778,3], joints [B,16,3]) in float64."""
    betas = np.asarray(betas, dtype=np.float64)
    B = betas.shape[0]
    go = np.zeros((B, 3)) if global_orient is None else np.asarray(global_orient, dtype=np.float64)
    hp = np.zeros((B, num_pca_comps if use_pca else N_POSE)) if hand_pose is None else np.asarray(hand_pose, dtype=np.float64)
    tr = np.zeros((B, 3)) if transl is None else np.asarray(transl, dtype=np.float64)
    if use_pca:
        hp = hp @ model["hands_components"][:num_pca_comps]
    pose_mean = np.concatenate([np.zeros(3), np.zeros(N_POSE) if flat_hand_mean else model["hands_mean"]])
    full_pose = np.concatenate([go, hp], axis=1) + pose_mean[None]
    v_shaped = model["v_template"][None] + np.einsum("bl,mkl->bmk", betas, model["shapedirs"])
    J = np.einsum("bik,ji->bjk", v_shaped, model["J_regressor"])
    R = rodrigues(full_pose.reshape(-1, 3)).reshape(B, N_JOINTS, 3, 3)
    pose_feature = (R[:, 1:] - np.eye(3)[None, None]).reshape(B, -1)
    posedirs = model["posedirs"].reshape(-1, 9 * (N_JOINTS - 1)).T            # [135, 778*3]
    v_posed = v_shaped + (pose_feature @ posedirs).reshape(B, N_VERTS, 3)
    parents = model["parents"]
    G = np.zeros((B, N_JOINTS, 4, 4))
    for j in range(N_JOINTS):
        T = np.zeros((B, 4, 4))
        T[:, :3, :3] = R[:, j]
        T[:, :3, 3] = J[:, j] if j == 0 else J[:, j] - J[:, parents[j]]
        T[:, 3, 3] = 1.0
        G[:, j] = T if j == 0 else G[:, parents[j]] @ T
    posed_joints = G[:, :, :3, 3].copy()
    A = G.copy()
    A[:, :, :3, 3] -= np.einsum("bjik,bjk->bji", G[:, :, :3, :3], J)
    T = np.einsum("vj,bjik->bvik", model["weights"], A)
    verts = np.einsum("bvik,bvk->bvi", T[:, :, :3, :3], v_posed) + T[:, :, :3, 3]
    return verts + tr[:, None], posed_joints + tr[:, None]
