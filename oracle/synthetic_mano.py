"""Synthetic MANO assets (TEST INFRASTRUCTURE — see oracle/README.md).

The real MANO_{LEFT,RIGHT}.pkl files are licence-gated and absent. The reference
reads these fields in manopth/manopth/manolayer.py:65-108 (through
mano/webuser/smpl_handpca_wrapper_HAND_only.py:22-67): hands_components (45,45),
hands_mean (45,), betas (10,), shapedirs (778,3,10), posedirs (778,3,135),
v_template (778,3), J_regressor sparse (16,778), weights (778,16), f (1538,3),
kintree_table (2,16).  We synthesise arrays of the same shape and the same
statistical role (hand-sized template in metres, convex skinning weights,
convex joint regressor) from a fixed numpy seed, so that every machine gets
bit-identical assets.
"""
import numpy as np

KINTREE_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
N_VERTS = 778
N_FACES = 1538


def make_mano_arrays(side: str) -> dict:
    """Return a dict of float32/int numpy arrays with the MANO schema."""
    assert side in ("left", "right")
    seed = 1 if side == "left" else 2
    rng = np.random.RandomState(seed)
    sign = -1.0 if side == "left" else 1.0

    # A crude hand: 16 joint centres (wrist + 5 fingers x 3) in metres, x mirrored per side.
    joints = np.zeros((16, 3), np.float64)
    finger_dirs = np.array([[-0.6, 0.7, 0.3], [-0.25, 1.0, 0.05], [0.0, 1.0, 0.0],
                            [0.25, 0.95, -0.03], [0.5, 0.8, -0.08]])
    finger_dirs /= np.linalg.norm(finger_dirs, axis=1, keepdims=True)
    # MANO joint order: index(1-3), middle(4-6), pinky(7-9), ring(10-12), thumb(13-15)
    order = [1, 2, 4, 3, 0]
    for f in range(5):
        d = finger_dirs[order[f]]
        base = 0.09 if order[f] != 0 else 0.035
        for k in range(3):
            joints[1 + 3 * f + k] = d * (base + 0.028 * k)
    joints[:, 0] *= sign

    # Vertices: each assigned to a "home" joint, scattered around it.
    home = rng.randint(0, 16, size=N_VERTS)
    home[:16] = np.arange(16)  # every joint owns at least one vertex
    v_template = joints[home] + rng.normal(0, 0.008, size=(N_VERTS, 3))

    # Skinning weights: convex, concentrated on home joint and its parent.
    w = rng.uniform(0, 0.05, size=(N_VERTS, 16))
    w[np.arange(N_VERTS), home] += 1.0
    par = np.array([max(p, 0) for p in KINTREE_PARENTS])[home]
    w[np.arange(N_VERTS), par] += rng.uniform(0, 0.6, size=N_VERTS)
    weights = w / w.sum(1, keepdims=True)

    # Joint regressor: convex combination of vertices homed at that joint (+ sparse noise).
    jr = np.zeros((16, N_VERTS))
    for j in range(16):
        idx = np.nonzero(home == j)[0]
        jr[j, idx] = rng.uniform(0.5, 1.0, size=idx.size)
        extra = rng.choice(N_VERTS, 12, replace=False)
        jr[j, extra] += rng.uniform(0, 0.05, size=12)
    jr /= jr.sum(1, keepdims=True)

    shapedirs = rng.normal(0, 0.004, size=(N_VERTS, 3, 10))
    if side == "left":
        # real MANO_LEFT ships shapedirs[:,0,:] un-mirrored ("shapedirs bug");
        # models/dir.py:306-309 flips it iff L~=R. Keep L != R here so that
        # branch is a no-op and both layers are used exactly as constructed.
        pass
    posedirs = rng.normal(0, 0.002, size=(N_VERTS, 3, 135))
    comps = rng.normal(0, 0.35, size=(45, 45))
    hands_mean = rng.normal(0, 0.25, size=(45,))
    betas = np.zeros((10,))
    faces = rng.randint(0, N_VERTS, size=(N_FACES, 3)).astype(np.int64)
    kintree = np.stack([np.array([4294967295] + KINTREE_PARENTS[1:], dtype=np.int64),
                        np.arange(16, dtype=np.int64)])
    return {
        "hands_components": comps.astype(np.float32),
        "hands_mean": hands_mean.astype(np.float32),
        "betas": betas.astype(np.float32),
        "shapedirs": shapedirs.astype(np.float32),
        "posedirs": posedirs.astype(np.float32),
        "v_template": v_template.astype(np.float32),
        "J_regressor": jr.astype(np.float32),
        "weights": weights.astype(np.float32),
        "f": faces,
        "kintree_table": kintree,
    }


def mano_buffers(side: str) -> dict:
    """The registered buffers of manopth ManoLayer (manolayer.py:71-101) as float32 numpy
    arrays, keyed by buffer name (th_*). These are what lives in the reference state_dict."""
    a = make_mano_arrays(side)
    return {
        "th_betas": a["betas"][None, :],
        "th_shapedirs": a["shapedirs"],
        "th_posedirs": a["posedirs"],
        "th_v_template": a["v_template"][None],
        "th_J_regressor": a["J_regressor"],
        "th_weights": a["weights"],
        "th_faces": a["f"],
        "th_hands_mean": a["hands_mean"][None, :],
        "th_comps": a["hands_components"],
        "th_selected_comps": a["hands_components"][:45],
    }
