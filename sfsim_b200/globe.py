"""Host-side mirror of sfsim.globe (src/clj/sfsim/globe.clj): `make_cube_map` writes the five files of every cube-map
tile exactly where and how the reference writes them; the pixels come from the GPU (sfsim_b200.cubemap.World over
libsfsim_atmosphere.so), the file formats from the standard library (gzip, tarfile) and Pillow (the reference uses STB
through LWJGL for JPEG and PNG).

    data/globe/<face>/<level>/<a>/<b>.jpg         day colours    spit-jpg        image.clj:109-120
    data/globe/<face>/<level>/<a>/<b>.night.jpg   night colours
    data/globe/<face>/<level>/<a>/<b>.water.gz    water bytes    spit-bytes-gz   util.clj:158-163  (row pitch align-address ct 4)
    data/globe/<face>/<level>/<a>/<b>.surf.gz     surface floats spit-floats-gz  util.clj:243-247  (little-endian float32)
    data/globe/<face>/<level>/<a>/<b>.png         normals        spit-normals    image.clj:126-136 (RGB8)
and `make_cube_map_tars` packs every data/globe/<face>/<level>/<a>/ into <a>.tar (globe.clj:83-96).
"""
import gzip
import os
import shutil
import tarfile

import numpy as np


def cube_path(prefix, face, level, y, x, suffix):
    """util.clj:300-304 (t_util.clj:123: (cube-path "globe" face5 2 3 1 ".png") => "globe/5/2/1/3.png")"""
    return "%s/%d/%d/%d/%d%s" % (prefix, face, level, x, y, suffix)


def cube_dir(prefix, face, level, x):
    """util.clj:307-311"""
    return "%s/%d/%d/%d" % (prefix, face, level, x)


def cube_tar(prefix, face, level, x):
    """util.clj:314-318"""
    return "%s/%d/%d/%d.tar" % (prefix, face, level, x)


def tile_path(prefix, level, y, x, suffix):
    """util.clj:286-290 (t_util.clj:115: (tile-path "world" 1 3 2 ".png") => "world/1/2/3.png")"""
    return "%s/%d/%d/%d%s" % (prefix, level, x, y, suffix)


def spit_bytes_gz(file_name, data):
    """util.clj:158-163"""
    with gzip.open(file_name, "wb") as f:
        f.write(np.ascontiguousarray(data, dtype=np.uint8).tobytes())


def slurp_bytes_gz(file_name):
    """util.clj:146-155"""
    with gzip.open(file_name, "rb") as f:
        return np.frombuffer(f.read(), dtype=np.uint8)


def spit_floats_gz(file_name, data):
    """util.clj:243-247: little-endian float32, gzip"""
    with gzip.open(file_name, "wb") as f:
        f.write(np.ascontiguousarray(data, dtype="<f4").tobytes())


def slurp_floats_gz(file_name):
    with gzip.open(file_name, "rb") as f:
        return np.frombuffer(f.read(), dtype="<f4")


def spit_jpg(path, rgba):
    """image.clj:109-120: an RGBA image as JPEG (STB drops the alpha channel)"""
    from PIL import Image
    Image.fromarray(np.ascontiguousarray(rgba, dtype=np.uint8)[..., :3], "RGB").save(path, "JPEG", quality=90)


def spit_normals(path, normals=None, normal_bytes=None):
    """image.clj:126-136: round(x 127.5 - 0.5) as a signed byte, written as the RGB8 bytes of a PNG.  `normal_bytes` is
    that array as the library delivers it (then `normals` is not needed)."""
    from PIL import Image
    if normal_bytes is None:
        scaled = np.asarray(normals, dtype=np.float32).astype(np.float64) * 127.5
        normal_bytes = np.floor((scaled - 0.5) + 0.5).astype(np.int8)
    Image.fromarray(np.ascontiguousarray(normal_bytes, dtype=np.int8).view(np.uint8), "RGB").save(path, "PNG")


def slurp_normals(path):
    """image.clj:139-156: (byte + 0.5) / 127.5 with the PNG's bytes read as signed"""
    from PIL import Image
    data = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8).view(np.int8)
    return ((data.astype(np.float64) + 0.5) / 127.5).astype(np.float32)


def write_cube_map_tile(prefix, face, level, b, a, tile):
    """globe.clj:73-78 for one tile dict (day, night, water, surface and normals or normal_bytes)"""
    os.makedirs(cube_dir(prefix, face, level, a), exist_ok=True)
    spit_jpg(cube_path(prefix, face, level, b, a, ".jpg"), tile["day"])
    spit_jpg(cube_path(prefix, face, level, b, a, ".night.jpg"), tile["night"])
    spit_bytes_gz(cube_path(prefix, face, level, b, a, ".water.gz"), tile["water"])
    spit_floats_gz(cube_path(prefix, face, level, b, a, ".surf.gz"), tile["surface"])
    spit_normals(cube_path(prefix, face, level, b, a, ".png"), tile.get("normals"), tile.get("normal_bytes"))


def make_cube_map(world, in_level, out_level, prefix="data/globe", rank=0, world_size=1, batch=256, **kw):
    """Program to generate tiles for cube map (globe.clj:29-80): the tiles of rank `rank` of `world_size`.  The float
    normals stay on the device: the PNG is written from the bytes the library encodes (44 % less PCIe traffic)."""
    outputs = ("day", "night", "water", "surface", "normal_bytes")
    return world.make_cube_map(in_level, out_level,
                               lambda key, tile: write_cube_map_tile(prefix, key[0], out_level, key[1], key[2], tile),
                               rank=rank, world_size=world_size, batch=batch, outputs=outputs, **kw)


def make_cube_map_tars(out_level, prefix="data/globe"):
    """Program to put cube map tiles into tar files (globe.clj:83-96)"""
    n = 1 << out_level
    for face in range(6):
        for a in range(n):
            directory = cube_dir(prefix, face, out_level, a)
            if not os.path.isdir(directory):
                continue
            with tarfile.open(cube_tar(prefix, face, out_level, a), "w") as tar:
                for name in sorted(os.listdir(directory)):
                    tar.add(os.path.join(directory, name), arcname=name)
            shutil.rmtree(directory)
