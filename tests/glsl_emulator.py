"""CPU emulation of the reference's GLSL LUT consumers, in float64 (test infrastructure).

Follows resources/shaders/atmosphere/{transmittance-track,transmittance-outer,ray-scatter-track,
ray-scatter-outer,phase-function}.glsl and resources/shaders/core/{transmittance-forward,
ray-scatter-forward,interpolate-2d,interpolate-4d,make-2d-index-from-4d,convert-2d-index,
elevation-to-index,height-to-index,sun-elevation-to-index,sun-angle-to-index,horizon-distance,
limit-quot,is-above-horizon}.glsl of wedesoft/sfsim.  Textures are the float32 arrays exactly as
the `.scatter` files hold them (RGB32F, GL_LINEAR, clamp-to-edge, texture.clj:355-404).
"""
import math

import numpy as np


def _texture(tex, u, v):
    """GL_LINEAR / CLAMP_TO_EDGE sample of tex[height][width][3] at normalised (u, v)."""
    h, w = tex.shape[:2]
    x = u * w - 0.5
    y = v * h - 0.5
    x0 = math.floor(x)
    y0 = math.floor(y)
    fx = x - x0
    fy = y - y0

    def px(ix, iy):
        return tex[min(max(iy, 0), h - 1), min(max(ix, 0), w - 1)].astype(np.float64)

    return ((px(x0, y0) * (1 - fx) + px(x0 + 1, y0) * fx) * (1 - fy) +
            (px(x0, y0 + 1) * (1 - fx) + px(x0 + 1, y0 + 1) * fx) * fy)


def limit_quot(a, b, lower, upper):
    if a == 0.0:
        return 0.0
    if b < 0:
        a, b = -a, -b
    if a < b * upper:
        return a / b if a > b * lower else lower
    return upper


def horizon_distance(ground_radius, radius_sqr):
    return math.sqrt(max(0.0, radius_sqr - ground_radius * ground_radius))


def height_to_index(radius, max_height, point):
    top = radius + max_height
    return horizon_distance(radius, float(np.dot(point, point))) / horizon_distance(radius, top * top)


def elevation_to_index(radius, max_height, point, direction, above_horizon):
    point_radius = float(np.linalg.norm(point))
    sin_elevation = float(np.dot(point, direction)) / point_radius
    rho = horizon_distance(radius, point_radius * point_radius)
    delta = point_radius * point_radius * sin_elevation * sin_elevation - rho * rho
    if above_horizon:
        top = radius + max_height
        h = math.sqrt(top * top - radius * radius)
        return 0.5 - limit_quot(point_radius * sin_elevation - math.sqrt(max(0.0, delta + h * h)), 2 * rho + 2 * h,
                                -0.5, 0.0)
    return 0.5 + limit_quot(point_radius * sin_elevation + math.sqrt(max(0.0, delta)), 2 * rho, -0.5, 0.0)


def sun_elevation_to_index(point, light):
    sin_elevation = float(np.dot(point, light)) / float(np.linalg.norm(point))
    return max(0.0, (1 - math.exp(-3 * sin_elevation - 0.6)) / (1 - math.exp(-3.6)))


def sun_angle_to_index(direction, light):
    return 0.5 * (1 + float(np.dot(direction, light)))


def is_above_horizon(radius, point, direction):
    dist = float(np.linalg.norm(point))
    s = float(np.dot(direction, point))
    return s >= 0 or s * s <= dist * dist - radius * radius


def phase(g, mu):
    return 3 * (1 - g * g) * (1 + mu * mu) / (8 * math.pi * (2 + g * g) * math.pow(1 + g * g - 2 * g * mu, 1.5))


class Atmosphere:
    """Holds the textures and uniforms the shaders read."""

    def __init__(self, radius, max_height, transmittance, ray_scatter=None, mie_strength=None, shape4=None):
        self.radius = radius
        self.max_height = max_height
        self.transmittance = transmittance          # [height][elevation][3] float32
        self.ray_scatter = ray_scatter              # [h*s][e*a][3] float32 (convert-4d-to-2d layout)
        self.mie_strength = mie_strength
        self.shape4 = shape4                        # (height, elevation, light-elevation, heading)

    # core/interpolate-2d.glsl + convert-2d-index.glsl
    def interpolate_2d(self, table, idx):
        size_y, size_x = table.shape[:2]
        px = idx[0] * (size_x - 1)
        py = idx[1] * (size_y - 1)
        return _texture(table, (px + 0.5) / size_x, (py + 0.5) / size_y)

    # core/interpolate-4d.glsl + make-2d-index-from-4d.glsl
    def interpolate_4d(self, table, idx):
        size_w, size_z, size_y, size_x = self.shape4
        pixel = [idx[0] * (size_x - 1), idx[1] * (size_y - 1), idx[2] * (size_z - 1), idx[3] * (size_w - 1)]
        frac_s = pixel[2] - math.floor(pixel[2])
        frac_t = pixel[3] - math.floor(pixel[3])
        z_floor = math.floor(pixel[2])
        w_floor = math.floor(pixel[3])
        div_x = size_x * size_z
        div_y = size_y * size_w
        s = (0.5 + pixel[0] + z_floor * size_x) / div_x
        t = (0.5 + pixel[0] + min(z_floor + 1, size_z - 1) * size_x) / div_x
        p = (0.5 + pixel[1] + w_floor * size_y) / div_y
        q = (0.5 + pixel[1] + min(w_floor + 1, size_w - 1) * size_y) / div_y
        return (_texture(table, s, p) * (1 - frac_s) * (1 - frac_t) + _texture(table, t, p) * frac_s * (1 - frac_t) +
                _texture(table, s, q) * (1 - frac_s) * frac_t + _texture(table, t, q) * frac_s * frac_t)

    def transmittance_forward(self, point, direction, above):
        return (elevation_to_index(self.radius, self.max_height, point, direction, above),
                height_to_index(self.radius, self.max_height, point))

    def ray_scatter_forward(self, point, direction, light, above):
        return (sun_angle_to_index(direction, light), sun_elevation_to_index(point, light),
                elevation_to_index(self.radius, self.max_height, point, direction, above),
                height_to_index(self.radius, self.max_height, point))

    # atmosphere/transmittance-outer.glsl
    def transmittance_outer(self, point, direction):
        point = np.asarray(point, dtype=np.float64)
        direction = np.asarray(direction, dtype=np.float64)
        return self.interpolate_2d(self.transmittance, self.transmittance_forward(point, direction, True))

    # atmosphere/transmittance-track.glsl
    def transmittance_track(self, p, q):
        p = np.asarray(p, dtype=np.float64)
        q = np.asarray(q, dtype=np.float64)
        dist = float(np.linalg.norm(q - p))
        if dist > 0:
            direction = (q - p) / dist
            above = is_above_horizon(self.radius, p, direction)
            t1 = self.interpolate_2d(self.transmittance, self.transmittance_forward(p, direction, above))
            t2 = self.interpolate_2d(self.transmittance, self.transmittance_forward(q, direction, above))
            return t1 / t2
        return np.ones(3)

    def _scatter(self, point, direction, light, above, mu):
        idx = self.ray_scatter_forward(point, direction, light, above)
        return self.interpolate_4d(self.ray_scatter, idx) + self.interpolate_4d(self.mie_strength, idx) * phase(0.76, mu)

    # atmosphere/ray-scatter-outer.glsl
    def ray_scatter_outer(self, light, point, direction):
        light = np.asarray(light, dtype=np.float64)
        point = np.asarray(point, dtype=np.float64)
        direction = np.asarray(direction, dtype=np.float64)
        return self._scatter(point, direction, light, True, float(np.dot(direction, light)))

    # atmosphere/ray-scatter-track.glsl
    def ray_scatter_track(self, light, p, q):
        light = np.asarray(light, dtype=np.float64)
        p = np.asarray(p, dtype=np.float64)
        q = np.asarray(q, dtype=np.float64)
        dist = float(np.linalg.norm(q - p))
        if dist > 0:
            direction = (q - p) / dist
            above = is_above_horizon(self.radius, p, direction)
            mu = float(np.dot(direction, light))
            sp = self._scatter(p, direction, light, above, mu)
            sq = self._scatter(q, direction, light, above, mu)
            return sp - self.transmittance_track(p, q) * sq
        return np.zeros(3)
