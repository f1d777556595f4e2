"""The offline build tasks of sfsim's build.clj that this library accelerates, under the names and with the output
files of the reference (`clj -T:build <task>`):

    worley          build.clj:34-38   data/clouds/worley-{north,south,cover}.raw
    perlin          build.clj:40-43   data/clouds/perlin.raw
    bluenoise       build.clj:45-51   data/bluenoise.raw
    atmosphere_lut  build.clj:84-87   data/atmosphere/{transmittance,surface-radiance,ray-scatter,mie-strength}.scatter
    cube_map        build.clj:294-298 data/globe/<face>/<level>/<a>.tar          (one level)
    cube_maps       build.clj:300-310 the pyramid of eight levels

    python -m sfsim_b200.build atmosphere_lut
"""
import os
import sys

import numpy as np

from . import atmosphere_lut as _al
from . import bluenoise as _bn
from . import globe as _globe
from . import perlin as _perlin
from . import worley as _worley


def _spit_floats(path, data):
    """util.clj:227-240: headerless little-endian float32"""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    np.ascontiguousarray(data, dtype="<f4").tofile(path)


def worley(size=16, divisions=4, out_dir="data/clouds"):
    """Generate 3D Worley noise textures"""
    for filename in ("worley-north.raw", "worley-south.raw", "worley-cover.raw"):
        _spit_floats(os.path.join(out_dir, filename), _worley.worley_noise(divisions, size))


def perlin(size=16, divisions=4, out_dir="data/clouds"):
    """Generate 3D Perlin noise textures"""
    _spit_floats(os.path.join(out_dir, "perlin.raw"), _perlin.perlin_noise(divisions, size))


def bluenoise(size=_bn.noise_size, out_dir="data"):
    """Generate 2D blue noise texture: n = size^2 / 10 seed samples, sigma 1.5"""
    _spit_floats(os.path.join(out_dir, "bluenoise.raw"), _bn.blue_noise_texture(size, (size * size) // 10, 1.5))


def atmosphere_lut(out_dir="data/atmosphere", num_gpus=1):
    """Generate atmospheric lookup tables"""
    return _al.generate_atmosphere_luts(out_dir, num_gpus=num_gpus)


def cube_map(world, in_level, out_level, prefix="data/globe"):
    """Create cube map level from map and elevation tiles held by `world` (sfsim_b200.cubemap.World)"""
    _globe.make_cube_map(world, in_level, out_level, prefix=prefix)
    _globe.make_cube_map_tars(out_level, prefix=prefix)


def cube_maps(world, prefix="data/globe"):
    """Create pyramid of cube maps"""
    for out_level in range(8):
        cube_map(world, out_level - 3, out_level, prefix=prefix)


if __name__ == "__main__":
    tasks = {"worley": worley, "perlin": perlin, "bluenoise": bluenoise, "atmosphere_lut": atmosphere_lut}
    if len(sys.argv) != 2 or sys.argv[1] not in tasks:
        raise SystemExit("usage: python -m sfsim_b200.build {%s}" % "|".join(tasks))
    tasks[sys.argv[1]]()
