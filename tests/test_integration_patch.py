"""integration/sfsim.patch must apply to the reference checkout (build.clj:84-87,294-298, deps.edn:42, Makefile:9-18,
scripts/packr-config-linux.json:7 plus the new shims src/clj/sfsim/atmosphere_cuda.clj and globe_cuda.clj).  The patched Clojure is not
executed anywhere in this repository: neither the build container nor the GPU box has a JVM."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
PATCH = os.path.join(ROOT, "integration", "sfsim.patch")
FILES = ["build.clj", "deps.edn", "Makefile", "scripts/packr-config-linux.json"]


def test_patch_is_committed_and_carries_the_shim():
    text = open(PATCH).read()
    for f in FILES + ["src/clj/sfsim/atmosphere_cuda.clj", "src/clj/sfsim/globe_cuda.clj"]:
        assert "+++ b/%s" % f in text
    for name in ("atmosphere_cuda.clj", "globe_cuda.clj"):
        shim = open(os.path.join(ROOT, "integration", "clj", "sfsim", name)).read()
        section = text.split("+++ b/src/clj/sfsim/%s" % name)[1].split("\ndiff --git")[0]
        added = "\n".join(line[1:] for line in section.splitlines() if line.startswith("+"))
        assert added.strip() == shim.strip()          # the patch ships exactly the shims kept in integration/clj
    assert "--enable-native-access=ALL-UNNAMED" in text and "libsfsim_atmosphere.so" in text


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference checkout is only present in the build container")
def test_patch_applies_to_the_reference(tmp_path):
    for f in FILES:
        os.makedirs(os.path.dirname(tmp_path / f), exist_ok=True)
        shutil.copy(os.path.join(REFERENCE, f), tmp_path / f)
    subprocess.check_call(["git", "init", "-q", "."], cwd=tmp_path)
    subprocess.check_call(["git", "apply", "--check", PATCH], cwd=tmp_path)
    subprocess.check_call(["git", "apply", PATCH], cwd=tmp_path)
    build = open(tmp_path / "build.clj").read()
    assert "requiring-resolve 'sfsim.atmosphere-cuda/generate-atmosphere-luts" in build
    assert "(al/generate-atmosphere-luts)" in build     # the CPU path stays the default
    assert os.path.exists(tmp_path / "src" / "clj" / "sfsim" / "atmosphere_cuda.clj")
    deps = open(tmp_path / "deps.edn").read()
    build_alias = deps.split(":build {")[1].split(":test")[0]
    assert "--enable-native-access=ALL-UNNAMED" in build_alias
    assert "libsfsim_atmosphere.so" in open(tmp_path / "scripts" / "packr-config-linux.json").read()
    assert "requiring-resolve 'sfsim.globe-cuda/make-cube-map" in build
    assert "(g/make-cube-map in-level out-level)" in build    # the CPU path stays the default
    assert os.path.exists(tmp_path / "src" / "clj" / "sfsim" / "globe_cuda.clj")
    makefile = open(tmp_path / "Makefile").read()
    assert "atm_lookup.cu" in makefile and "cubemap.cu" in makefile
