"""Host-side scene API: the reference's primitive structs and factories, in numpy.

Mirrors, name for name and default for default:
  * the POD structs of src/scene.h:7-154 as numpy structured dtypes whose field
    offsets equal the std140 layout the shader reads (rt.frag:24-113; table in
    include/rtb200_types.h) — an array of `rt_sphere` can be handed to
    GLWrapper.init_buffer / rtb_upload byte-for-byte;
  * the static factories of SceneManager (src/SceneManager.h:17-25,
    src/SceneManager.cpp:137-236);
  * SurfaceFactory (src/Surface.h:7-97);
  * scene_container.get_defines (src/scene.h:128-153).
The C++ host (raytracing-opengl_b200/host/) reuses the reference's own
SceneManager unchanged; this module is the same API for Python callers
(tests, bench.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)
f4, i4, u4 = np.float32, np.int32, np.uint32


def _dt(fields, size):
    names, formats, offsets = zip(*fields)
    return np.dtype({"names": list(names), "formats": list(formats), "offsets": list(offsets), "itemsize": size})


rt_material = _dt([("color", (f4, 3), 0), ("absorb", (f4, 3), 16), ("diffuse", f4, 28), ("reflect", f4, 32),
                   ("refract", f4, 36), ("specular", i4, 40), ("kd", f4, 44), ("ks", f4, 48)], 64)
rt_sphere = _dt([("material", rt_material, 0), ("obj", (f4, 4), 64), ("quat_rotation", (f4, 4), 80),
                 ("textureNum", i4, 96), ("hollow", u4, 100)], 112)
rt_plane = _dt([("material", rt_material, 0), ("pos", (f4, 3), 64), ("normal", (f4, 3), 80)], 96)
rt_box = _dt([("mat", rt_material, 0), ("quat_rotation", (f4, 4), 64), ("pos", (f4, 3), 80), ("form", (f4, 3), 96),
              ("textureNum", i4, 108)], 112)
rt_torus = _dt([("mat", rt_material, 0), ("quat_rotation", (f4, 4), 64), ("pos", (f4, 3), 80), ("form", (f4, 2), 96)], 112)
rt_ring = _dt([("mat", rt_material, 0), ("quat_rotation", (f4, 4), 64), ("pos", (f4, 3), 80), ("textureNum", i4, 92),
               ("r1", f4, 96), ("r2", f4, 100)], 112)
rt_surface = _dt([("mat", rt_material, 0), ("quat_rotation", (f4, 4), 64), ("v_min", (f4, 3), 80), ("v_max", (f4, 3), 96),
                  ("pos", (f4, 3), 112), ("a", f4, 124), ("b", f4, 128), ("c", f4, 132), ("d", f4, 136), ("e", f4, 140),
                  ("f", f4, 144)], 160)
rt_light_direct = _dt([("direction", (f4, 3), 0), ("color", (f4, 3), 16), ("intensity", f4, 28)], 32)
rt_light_point = _dt([("pos", (f4, 4), 0), ("color", (f4, 3), 16), ("intensity", f4, 28), ("linear_k", f4, 32),
                      ("quadratic_k", f4, 36)], 48)
rt_scene = _dt([("quat_camera_rotation", (f4, 4), 0), ("camera_pos", (f4, 3), 16), ("bg_color", (f4, 3), 32),
                ("canvas_width", i4, 44), ("canvas_height", i4, 48), ("reflect_depth", i4, 52)], 64)
rt_defines = np.dtype([("sphere_size", i4), ("plane_size", i4), ("surface_size", i4), ("box_size", i4), ("torus_size", i4),
                       ("ring_size", i4), ("light_point_size", i4), ("light_direct_size", i4), ("iterations", i4),
                       ("ambient_color", f4, 3), ("shadow_ambient", f4, 3)])

QUAT_IDENTITY = (0.0, 0.0, 0.0, 1.0)      # x,y,z,w: glm::quat(1,0,0,0) in memory (scene.h:40)

# uniform-block name -> (binding point, element dtype); SceneManager.cpp:246-254
UBO_BLOCKS = {
    "scene_buf": (0, rt_scene), "spheres_buf": (1, rt_sphere), "planes_buf": (2, rt_plane), "surfaces_buf": (3, rt_surface),
    "boxes_buf": (4, rt_box), "toruses_buf": (5, rt_torus), "rings_buf": (6, rt_ring),
    "lights_point_buf": (7, rt_light_point), "lights_direct_buf": (8, rt_light_direct),
}


def _one(dtype):
    return np.zeros((), dtype=dtype)     # `= {}` value-initialisation of the factories


# ---------------------------------------------------------------------------
# quaternion helpers (glm semantics, x,y,z,w storage, fp32 results)
# ---------------------------------------------------------------------------
def quat_from_euler(pitch, yaw, roll):
    """glm::quat(glm::vec3 eulerAngles) (glm/detail/type_quat.inl): returns x,y,z,w."""
    cx, cy, cz = (f4(math.cos(f4(a) * f4(0.5))) for a in (pitch, yaw, roll))
    sx, sy, sz = (f4(math.sin(f4(a) * f4(0.5))) for a in (pitch, yaw, roll))
    w = cx * cy * cz + sx * sy * sz
    x = sx * cy * cz - cx * sy * sz
    y = cx * sy * cz + sx * cy * sz
    z = cx * cy * sz - sx * sy * cz
    return np.array([x, y, z, w], dtype=f4)


def quat_angle_axis(angle, axis):
    """glm::angleAxis(angle, axis): x,y,z,w."""
    a = f4(angle)
    s = f4(math.sin(a * f4(0.5)))
    ax = np.asarray(axis, dtype=f4)
    return np.array([ax[0] * s, ax[1] * s, ax[2] * s, f4(math.cos(a * f4(0.5)))], dtype=f4)


def quat_mul(p, q):
    """glm operator*(quat p, quat q), x,y,z,w."""
    p = np.asarray(p, dtype=f4)
    q = np.asarray(q, dtype=f4)
    px, py, pz, pw = p
    qx, qy, qz, qw = q
    return np.array([pw * qx + px * qw + py * qz - pz * qy,
                     pw * qy + py * qw + pz * qx - px * qz,
                     pw * qz + pz * qw + px * qy - py * qx,
                     pw * qw - px * qx - py * qy - pz * qz], dtype=f4)


# ---------------------------------------------------------------------------
# SceneManager factories (src/SceneManager.cpp:137-236)
# ---------------------------------------------------------------------------
class SceneManager:
    @staticmethod
    def create_material(color, specular, reflect, refract=0.0, absorb=(0, 0, 0), diffuse=0.7, kd=0.8, ks=0.2):
        m = _one(rt_material)
        m["color"] = color
        m["absorb"] = absorb
        m["specular"] = int(specular)
        m["reflect"] = reflect
        m["refract"] = refract
        m["diffuse"] = diffuse
        m["kd"] = kd
        m["ks"] = ks
        return m

    @staticmethod
    def create_sphere(center, radius, material, hollow=False):
        s = _one(rt_sphere)
        s["obj"] = (*center, radius)
        s["hollow"] = 1 if hollow else 0
        s["material"] = material
        s["quat_rotation"] = QUAT_IDENTITY
        return s

    @staticmethod
    def create_plane(normal, pos, material):
        p = _one(rt_plane)
        p["normal"] = normal
        p["pos"] = pos
        p["material"] = material
        return p

    @staticmethod
    def create_box(pos, form, material):
        b = _one(rt_box)
        b["form"] = form
        b["pos"] = pos
        b["mat"] = material
        b["quat_rotation"] = QUAT_IDENTITY
        return b

    @staticmethod
    def create_torus(pos, form, material):
        t = _one(rt_torus)
        t["form"] = form
        t["pos"] = pos
        t["mat"] = material
        t["quat_rotation"] = QUAT_IDENTITY
        return t

    @staticmethod
    def create_ring(pos, r1, r2, material):
        r = _one(rt_ring)
        r["pos"] = pos
        r["mat"] = material
        r["r1"] = f4(r1) * f4(r1)       # squared radii, SceneManager.cpp:195-196
        r["r2"] = f4(r2) * f4(r2)
        r["quat_rotation"] = QUAT_IDENTITY
        return r

    @staticmethod
    def create_light_point(position, color, intensity, linear_k=0.22, quadratic_k=0.2):
        l = _one(rt_light_point)
        l["intensity"] = intensity
        l["pos"] = position
        l["color"] = color
        l["linear_k"] = linear_k
        l["quadratic_k"] = quadratic_k
        return l

    @staticmethod
    def create_light_direct(direction, color, intensity):
        l = _one(rt_light_direct)
        l["intensity"] = intensity
        l["direction"] = direction
        l["color"] = color
        return l

    @staticmethod
    def create_scene(width, height):
        s = _one(rt_scene)
        s["canvas_height"] = height
        s["canvas_width"] = width
        s["bg_color"] = (0, 0, 0)
        s["reflect_depth"] = 5          # SceneManager.cpp:233
        s["quat_camera_rotation"] = (0, 0, 0, 0)   # value-initialised glm::quat; set by SceneManager::update_scene
        return s


# ---------------------------------------------------------------------------
# SurfaceFactory (src/Surface.h)
# ---------------------------------------------------------------------------
def _surface(material, **coef):
    s = _one(rt_surface)
    s["quat_rotation"] = QUAT_IDENTITY
    s["v_min"] = (-FLT_MAX,) * 3
    s["v_max"] = (FLT_MAX,) * 3
    for k, v in coef.items():
        s[k] = v
    s["mat"] = material
    return s


def _inv2(x):
    return f4(math.pow(float(f4(x)), -2.0))   # powf(a, -2)


class SurfaceFactory:
    @staticmethod
    def GetEllipsoid(a, b, c, material):
        return _surface(material, a=_inv2(a), b=_inv2(b), c=_inv2(c), f=-1)

    @staticmethod
    def GetEllipticParaboloid(a, b, material):
        return _surface(material, a=_inv2(a), b=_inv2(b), d=-1)

    @staticmethod
    def GetHyperbolicParaboloid(a, b, material):
        return _surface(material, a=_inv2(a), b=-_inv2(b), d=-1)

    @staticmethod
    def GetEllipticHyperboloidOneSheet(a, b, c, material):
        return _surface(material, a=_inv2(a), b=_inv2(b), c=-_inv2(c), f=-1)

    @staticmethod
    def GetEllipticHyperboloidTwoSheets(a, b, c, material):
        return _surface(material, a=_inv2(a), b=_inv2(b), c=-_inv2(c), f=1)

    @staticmethod
    def GetEllipticCone(a, b, c, material):
        return _surface(material, a=_inv2(a), b=_inv2(b), c=-_inv2(c))

    @staticmethod
    def GetEllipticCylinder(a, b, material):
        return _surface(material, a=_inv2(a), b=_inv2(b), f=-1)

    @staticmethod
    def GetHyperbolicCylinder(a, b, material):
        return _surface(material, a=_inv2(a), b=-_inv2(b), f=-1)

    @staticmethod
    def GetParabolicCylinder(a, material):
        return _surface(material, a=1, e=2 * a)


# ---------------------------------------------------------------------------
# scene_container (src/scene.h:128-154)
# ---------------------------------------------------------------------------
@dataclass
class SceneContainer:
    scene: np.ndarray = field(default_factory=lambda: _one(rt_scene))
    ambient_color: tuple = (0.0, 0.0, 0.0)
    shadow_ambient: tuple = (0.0, 0.0, 0.0)
    spheres: list = field(default_factory=list)
    planes: list = field(default_factory=list)
    surfaces: list = field(default_factory=list)
    boxes: list = field(default_factory=list)
    toruses: list = field(default_factory=list)
    rings: list = field(default_factory=list)
    lights_point: list = field(default_factory=list)
    lights_direct: list = field(default_factory=list)

    _ARRAYS = (("spheres", rt_sphere), ("planes", rt_plane), ("surfaces", rt_surface), ("boxes", rt_box),
               ("toruses", rt_torus), ("rings", rt_ring), ("lights_point", rt_light_point), ("lights_direct", rt_light_direct))

    def array(self, name) -> np.ndarray:
        """Contiguous structured array of one primitive list (what std::vector<T>::data() holds)."""
        dt = dict(self._ARRAYS)[name]
        items = getattr(self, name)
        if isinstance(items, np.ndarray):
            return np.ascontiguousarray(items, dtype=dt)
        out = np.zeros(len(items), dtype=dt)
        for i, it in enumerate(items):
            out[i] = it
        return out

    def get_defines(self) -> np.ndarray:
        d = np.zeros((), dtype=rt_defines)
        d["sphere_size"] = len(self.spheres)
        d["plane_size"] = len(self.planes)
        d["surface_size"] = len(self.surfaces)
        d["box_size"] = len(self.boxes)
        d["torus_size"] = len(self.toruses)
        d["ring_size"] = len(self.rings)
        d["light_point_size"] = len(self.lights_point)
        d["light_direct_size"] = len(self.lights_direct)
        d["iterations"] = int(self.scene["reflect_depth"])
        d["ambient_color"] = self.ambient_color
        d["shadow_ambient"] = self.shadow_ambient
        return d

    def uses_textures(self) -> bool:
        return any(int(a["textureNum"]) != 0 for name in ("spheres", "boxes", "rings") for a in self.array(name))

    def scene_bytes(self) -> int:
        return 64 + sum(self.array(n).nbytes for n, _ in self._ARRAYS)
