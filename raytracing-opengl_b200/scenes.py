"""Scene scripts: the reference's default scene and the synthetic benchmark scenes.

`default_scene()` restates src/main.cpp:43-132 plus the first call of
update_scene (src/main.cpp:197-246) and SceneManager::update_scene
(src/SceneManager.cpp:43-74) at the deterministic clock value t = 0, dt = 0 —
the state the reference renders in its first frame.  (The C++ host in host/
runs the reference's own main.cpp instead; tests/golden/default_scene_t0.npz
is the byte dump of that run and pins this restatement.)

`synthetic_scene(name)` generates the BASELINE.json configs:
`spheres4k`, `tori1080`, `mixed1024`.  The generator below IS the definition of
those workloads: PCG32 (XSH-RR, stream 1), 24-bit uniforms, fixed draw order.
"""
from __future__ import annotations

import math

import numpy as np

from .scene import (FLT_MAX, SceneContainer, SceneManager, SurfaceFactory, f4, quat_angle_axis, quat_from_euler,
                    quat_mul)

# BASELINE.json configs -> (scene name, width, height, iterations)
CONFIGS = {
    "default256": ("default", 256, 256, 1),
    "default1080": ("default", 1920, 1080, 4),
    "spheres4k": ("spheres4k", 3840, 2160, 8),
    "tori1080": ("tori1080", 1920, 1080, 4),
    "mixed1024_4k": ("mixed1024", 3840, 2160, 8),
    "mixed1024_8k": ("mixed1024", 7680, 4320, 8),
}

# texture unit -> file under ASSETS_DIR/textures (main.cpp:149-153); 0 = skybox faces (main.cpp:137-145)
DEFAULT_TEXTURES = {1: "8k_jupiter.jpg", 2: "8k_saturn.jpg", 3: "2k_mars.jpg", 4: "8k_saturn_ring_alpha.png", 5: "container.png"}
DEFAULT_CUBEMAP = ["sb_nebula/GalaxyTex_PositiveX.jpg", "sb_nebula/GalaxyTex_NegativeX.jpg", "sb_nebula/GalaxyTex_PositiveY.jpg",
                   "sb_nebula/GalaxyTex_NegativeY.jpg", "sb_nebula/GalaxyTex_PositiveZ.jpg", "sb_nebula/GalaxyTex_NegativeZ.jpg"]


def _default_lights(scene: SceneContainer):
    scene.shadow_ambient = (0.1, 0.1, 0.1)          # main.cpp:47-48
    scene.ambient_color = (0.025, 0.025, 0.025)
    scene.lights_point.append(SceneManager.create_light_point((3, 5, 0, 0.1), (1, 1, 1), 25.5))     # main.cpp:51
    scene.lights_direct.append(SceneManager.create_light_direct((3, -1, 1), (1, 1, 1), 1.5))        # main.cpp:52


def default_scene(width=1280, height=720, iterations=None, textured=True) -> SceneContainer:
    """src/main.cpp:43-132 at t = 0.  `textured=False` clears every textureNum (no 2-D samplers needed)."""
    if width % 2:
        width += 1                                   # main.cpp:40-41
    if height % 2:
        height += 1
    sc = SceneContainer()
    sc.scene = SceneManager.create_scene(width, height)
    sc.scene["camera_pos"] = (0, 0, -5)
    _default_lights(sc)
    cm = SceneManager.create_material
    sc.spheres.append(SceneManager.create_sphere((2, 0, 6), 1, cm((0, 0, 1), 50, 0.35)))
    sc.spheres.append(SceneManager.create_sphere((-1, 0, 6), 1, cm((1, 0, 0), 100, 0.1), True))
    sc.spheres.append(SceneManager.create_sphere((0.5, 2, 6), 1, cm((1, 1, 1), 200, 0.1, 1.125, (1, 0, 2), 1), True))

    saturn_pitch = quat_from_euler(math.radians(15.0), 0, 0)
    jupiter = SceneManager.create_sphere((0, 0, 0), 5000, cm((0, 0, 0), 0, 0.0))
    jupiter["textureNum"] = 1
    saturn_radius = 4150
    saturn = SceneManager.create_sphere((0, 0, 0), saturn_radius, cm((0, 0, 0), 0, 0.0))
    saturn["textureNum"] = 2
    saturn["quat_rotation"] = saturn_pitch
    mars = SceneManager.create_sphere((0, 0, 0), 500, cm((0, 0, 0), 0, 0.0))
    mars["textureNum"] = 3
    ring = SceneManager.create_ring((0, 0, 0), saturn_radius * 1.1166, saturn_radius * 2.35, cm((0, 0, 0), 0, 0))
    ring["textureNum"] = 4
    ring["quat_rotation"] = quat_mul(quat_angle_axis(math.radians(90.0), (1, 0, 0)), saturn_pitch)

    # update_scene(scene, dt=0, t=0), main.cpp:197-232: positions only (all angleAxis(0) are identity)
    t = f4(0.0)
    jupiter["obj"][0] = math.cos(t * f4(0.02)) * 20000
    jupiter["obj"][2] = math.sin(t * f4(0.02)) * 20000
    sx, sz = math.cos(float(t * f4(0.0082) + f4(1))) * 35000, math.sin(float(t * f4(0.0082) + f4(1))) * 35000
    saturn["obj"][0], saturn["obj"][2] = sx, sz
    ring["pos"][0], ring["pos"][2] = sx, sz
    mars["obj"][0] = math.cos(float(t * f4(0.05) + f4(0.5))) * 10000
    mars["obj"][2] = math.sin(float(t * f4(0.05) + f4(0.5))) * 10000
    mars["obj"][1] = -math.cos(float(t * f4(0.05))) * 3000
    sc.spheres += [jupiter, saturn, mars]
    sc.rings.append(ring)

    sc.boxes.append(SceneManager.create_box((0, -1.2, 6), (10, 0.2, 5), cm((1, 0.6, 0), 100, 0.05)))
    box = SceneManager.create_box((8, 1, 6), (1, 1, 1), cm((0.8, 0.7, 0), 50, 0.0))
    box["textureNum"] = 5
    sc.boxes.append(box)

    torus = SceneManager.create_torus((-9, 0.5, 6), (1.0, 0.5), cm((0.5, 0.4, 1), 200, 0.2))
    torus["quat_rotation"] = quat_from_euler(math.radians(45.0), 0, 0)
    sc.toruses.append(torus)

    cone = SurfaceFactory.GetEllipticCone(1 / 3.0, 1 / 3.0, 1, cm((234 / 255.0, 17 / 255.0, 82 / 255.0), 200, 0.2))
    cone["pos"] = (-5, 4, 6)
    cone["quat_rotation"] = quat_from_euler(math.radians(90.0), 0, 0)
    cone["v_min"][1] = -1
    cone["v_max"][1] = 4
    sc.surfaces.append(cone)
    cyl = SurfaceFactory.GetEllipticCylinder(1 / 2.0, 1 / 2.0, cm((200 / 255.0, 255 / 255.0, 0 / 255.0), 200, 0.2))
    cyl["pos"] = (5, 0, 6)
    cyl["quat_rotation"] = quat_from_euler(math.radians(90.0), 0, 0)
    cyl["v_min"][1] = -1
    cyl["v_max"][1] = 1
    sc.surfaces.append(cyl)

    # SceneManager::update_scene, SceneManager.cpp:43-74 with yaw = pitch = 0
    sc.scene["quat_camera_rotation"] = quat_from_euler(-0.0, 0.0, 0)
    if iterations is not None:
        sc.scene["reflect_depth"] = iterations
    if not textured:
        for lst in (sc.spheres, sc.boxes, sc.rings):
            for p in lst:
                p["textureNum"] = 0
    return sc


# ---------------------------------------------------------------------------
# PCG32 (O'Neill, XSH-RR 64/32), stream 1
# ---------------------------------------------------------------------------
class PCG32:
    MULT = 6364136223846793005
    MASK = (1 << 64) - 1

    def __init__(self, seed: int, stream: int = 1):
        self.state = 0
        self.inc = ((stream << 1) | 1) & self.MASK
        self.next_u32()
        self.state = (self.state + seed) & self.MASK
        self.next_u32()

    def next_u32(self) -> int:
        old = self.state
        self.state = (old * self.MULT + self.inc) & self.MASK
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def uniform(self) -> np.float32:
        """24-bit uniform in [0,1), exactly representable in fp32."""
        return f4((self.next_u32() >> 8) * (1.0 / 16777216.0))

    def range(self, lo, hi) -> np.float32:
        return f4(f4(lo) + f4(f4(hi) - f4(lo)) * self.uniform())

    def index(self, n: int) -> int:
        return int(self.next_u32() % n)

    def quat(self) -> np.ndarray:
        """Uniform random unit quaternion: rejection-sample the 4-ball, normalise (x,y,z,w)."""
        while True:
            q = np.array([self.range(-1, 1) for _ in range(4)], dtype=np.float64)
            n2 = float(q @ q)
            if 1e-4 < n2 <= 1.0:
                return (q / math.sqrt(n2)).astype(f4)


SPECULARS = (0, 10, 50, 100, 200)


def _rand_material(rng: PCG32):
    color = (rng.range(0.1, 1), rng.range(0.1, 1), rng.range(0.1, 1))
    specular = SPECULARS[rng.index(5)]
    reflect = rng.range(0, 0.6)
    if rng.uniform() < 0.25:
        reflect = 0.0                      # 25 % purely diffuse
    return SceneManager.create_material(color, specular, reflect)


def _rand_centre(rng: PCG32):
    return (rng.range(-20, 20), rng.range(0.3, 10), rng.range(0, 40))


def _base(width, height, iterations) -> SceneContainer:
    sc = SceneContainer()
    sc.scene = SceneManager.create_scene(width, height)
    sc.scene["camera_pos"] = (0, 2, -12)
    sc.scene["quat_camera_rotation"] = (0, 0, 0, 1)
    sc.scene["reflect_depth"] = iterations
    _default_lights(sc)
    return sc


def _add_spheres(sc, rng, n):
    for _ in range(n):
        c = _rand_centre(rng)
        r = rng.range(0.3, 1.5)
        sc.spheres.append(SceneManager.create_sphere(c, r, _rand_material(rng)))


def _add_boxes(sc, rng, n):
    for _ in range(n):
        c = _rand_centre(rng)
        form = (rng.range(0.3, 1.5), rng.range(0.3, 1.5), rng.range(0.3, 1.5))
        q = rng.quat()
        b = SceneManager.create_box(c, form, _rand_material(rng))
        b["quat_rotation"] = q
        sc.boxes.append(b)


def _add_tori(sc, rng, n):
    for _ in range(n):
        c = _rand_centre(rng)
        form = (rng.range(0.6, 1.5), rng.range(0.15, 0.5))
        q = rng.quat()
        t = SceneManager.create_torus(c, form, _rand_material(rng))
        t["quat_rotation"] = q
        sc.toruses.append(t)


def _add_quadrics(sc, rng, n_each):
    for kind in ("ellipsoid", "cone", "cylinder"):
        for _ in range(n_each):
            c = _rand_centre(rng)
            a, b, cc = rng.range(0.3, 1.2), rng.range(0.3, 1.2), rng.range(0.3, 1.2)
            q = rng.quat()
            m = _rand_material(rng)
            if kind == "ellipsoid":
                s = SurfaceFactory.GetEllipsoid(a, b, cc, m)
            elif kind == "cone":
                s = SurfaceFactory.GetEllipticCone(a, b, cc, m)
            else:
                s = SurfaceFactory.GetEllipticCylinder(a, b, m)
            s["pos"] = c
            s["quat_rotation"] = q
            s["v_min"] = tuple(f4(x) - f4(2) for x in c)     # world-space clip box = centre +- 2
            s["v_max"] = tuple(f4(x) + f4(2) for x in c)
            sc.surfaces.append(s)


def synthetic_scene(name: str, width: int, height: int, iterations: int) -> SceneContainer:
    sc = _base(width, height, iterations)
    if name == "spheres4k":
        rng = PCG32(3)
        _add_spheres(sc, rng, 256)
        _add_boxes(sc, rng, 64)
        sc.planes.append(SceneManager.create_plane((0, 1, 0), (0, 0, 0), _rand_material(rng)))
    elif name == "tori1080":
        rng = PCG32(4)
        _add_tori(sc, rng, 128)
    elif name == "mixed1024":
        rng = PCG32(5)
        _add_spheres(sc, rng, 512)
        _add_boxes(sc, rng, 256)
        _add_quadrics(sc, rng, 64)
        _add_tori(sc, rng, 64)
    elif name.startswith("mini"):
        # small mixed scene for fast tests: mini<seed>
        rng = PCG32(int(name[4:] or 1))
        _add_spheres(sc, rng, 6)
        _add_boxes(sc, rng, 4)
        _add_quadrics(sc, rng, 1)
        _add_tori(sc, rng, 2)
        sc.planes.append(SceneManager.create_plane((0, 1, 0), (0, 0, 0), _rand_material(rng)))
    else:
        raise ValueError(f"unknown synthetic scene {name!r}")
    return sc


def build_config(config: str, scale: float = 1.0) -> SceneContainer:
    """Scene for one BASELINE.json config; `scale` shrinks the canvas (tests), keeping even sizes."""
    name, w, h, it = CONFIGS[config]
    w = max(2, int(round(w * scale)) & ~1)
    h = max(2, int(round(h * scale)) & ~1)
    if name == "default":
        return default_scene(w, h, it)
    return synthetic_scene(name, w, h, it)
