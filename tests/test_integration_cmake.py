"""The integration artefacts a vierkant maintainer would use: the CMake module configures (option VIERKANT_BCN_CUDA swaps
src/texture_block_compression.cpp for the CUDA drop-in and links libvierkant_bcn_cuda) and integration/vierkant.patch applies
to the reference tree.  CPU only; no compute."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_cmake_module_configures_and_swaps_the_translation_unit(tmp_path):
    cmake = shutil.which("cmake")
    if cmake is None or shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("cmake / nvcc not available")
    gen = ["-G", "Ninja"] if shutil.which("ninja") else []
    for option, want_cuda in (("ON", True), ("OFF", False)):
        build = tmp_path / option
        out = subprocess.run([cmake, *gen, "-S", os.path.join(ROOT, "integration", "cmake_check"), "-B", str(build),
                              f"-DVIERKANT_BCN_CUDA={option}"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
        sources = [l for l in out.stdout.splitlines() if "vierkant SOURCES:" in l][0]
        links = [l for l in out.stdout.splitlines() if "vierkant LINK_LIBRARIES:" in l][0]
        assert ("texture_block_compression_cuda.cpp" in sources) == want_cuda
        assert ("src/texture_block_compression.cpp" in sources) == (not want_cuda)
        assert ("vierkant_bcn_cuda" in links) == want_cuda
        assert "other.cpp" in sources


def test_vierkant_patch_applies_to_the_reference_tree(tmp_path):
    if not os.path.isdir(REF) or shutil.which("patch") is None:
        pytest.skip("reference tree or patch(1) not available")
    for rel in ("CMakeLists.txt", "src/CMakeLists.txt", "src/model/model_loading.cpp"):
        dst = tmp_path / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(REF, rel), dst)
    with open(os.path.join(ROOT, "integration", "vierkant.patch")) as f:
        out = subprocess.run(["patch", "-p1"], stdin=f, cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    text = (tmp_path / "src/model/model_loading.cpp").read_text()
    assert "bcn::compress(std::span<const bcn::compress_info_t>(compress_infos))" in text
    assert "vierkant_bcn_cuda_sources(FOLDER_SOURCES)" in (tmp_path / "src/CMakeLists.txt").read_text()
