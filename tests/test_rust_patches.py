"""The reference-side change set (rust/patches/*.diff, rust/src/gpu.rs, rust/build.rs — SURVEY §8 row f1).

There is no Rust toolchain in this image, so the patches cannot be compiled; what can be checked is that they are a real,
current patch set: every diff applies cleanly (`git apply --check`, then for real) to a scratch copy of the reference, the
result is exactly what rust/make_patches.py generates, and the shim's `#[repr(C)]` mirrors and `extern "C"` block agree
with include/ptgpu.h (struct sizes through the compiled library, every imported function declared in the header)."""
import glob
import os
import re
import shutil
import subprocess

import pytest

import pathtrace_rs_b200 as pt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCHES = sorted(glob.glob(os.path.join(ROOT, "rust", "patches", "*.diff")))


def test_patch_set_covers_the_files_the_drop_in_needs():
    touched = set()
    for p in PATCHES:
        for line in open(p):
            if line.startswith("+++ b/"):
                touched.add(line[6:].strip())
    assert touched == {"Cargo.toml", "src/bench.rs", "src/camera.rs", "src/collision/hitable_list.rs", "src/collision/moving_sphere.rs",
                       "src/main.rs", "src/params.rs", "src/perlin.rs", "src/scene.rs", "src/texture.rs"}


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the development container")
def test_patches_apply_to_a_scratch_copy_of_the_reference(tmp_path):
    work = tmp_path / "pathtrace-rs"
    work.mkdir()
    shutil.copy(os.path.join(REF, "Cargo.toml"), work)
    shutil.copytree(os.path.join(REF, "src"), work / "src")
    for p in PATCHES:
        subprocess.run(["git", "apply", "--check", "-p1", p], cwd=work, check=True)
    for p in PATCHES:
        subprocess.run(["git", "apply", "-p1", p], cwd=work, check=True)
    shutil.copy(os.path.join(ROOT, "rust", "build.rs"), work)
    shutil.copy(os.path.join(ROOT, "rust", "src", "gpu.rs"), work / "src")
    # the patched tree names everything the shim calls, and the call site of the drop-in is untouched
    src = {os.path.relpath(f, work): open(f).read() for f in glob.glob(str(work / "**" / "*.rs"), recursive=True)}
    shim = src["src/gpu.rs"]
    for accessor, where in (("to_ffi", "src/camera.rs"), ("world", "src/scene.rs"), ("sky_colour", "src/scene.rs"), ("attach_gpu", "src/scene.rs"),
                            ("hitables", "src/collision/hitable_list.rs"), ("endpoints", "src/collision/moving_sphere.rs"),
                            ("tables", "src/perlin.rs"), ("raw", "src/texture.rs")):
        assert re.search(r"pub\(crate\) fn %s\b" % accessor, src[where]), (accessor, where)
        if accessor != "attach_gpu":
            assert re.search(r"\.%s\(" % accessor, shim), accessor
    assert len(re.findall(r"fn sky\b", src["src/scene.rs"])) == 1  # no second `sky`: Rust has no overloading
    assert "scene.update(&params, &camera, frame_num, &mut rgb_buffer)" in src["src/offline.rs"]  # offline.rs:29 as upstream
    assert "scene.update(&params, &camera, frame_num, &mut rgb_buffer)" in src["src/glium_window.rs"]
    assert "gpu.update(params, camera, frame_num, buffer)" in src["src/scene.rs"]
    assert 'build = "build.rs"' in open(work / "Cargo.toml").read() and "gpu = []" in open(work / "Cargo.toml").read()
    # regenerating the patches from the reference gives the committed files
    regen = tmp_path / "regen"
    shutil.copytree(os.path.join(ROOT, "rust"), regen)
    subprocess.run(["python", str(regen / "make_patches.py"), REF], check=True, stdout=subprocess.DEVNULL)
    for p in PATCHES:
        assert open(p).read() == open(regen / "patches" / os.path.basename(p)).read(), os.path.basename(p)


def _rust_struct_size(body):
    """size of a #[repr(C)] struct made of u8/u32/i32/f32/u64/pointer fields and fixed arrays of them (natural alignment)"""
    sizes = {"u8": 1, "i32": 4, "u32": 4, "f32": 4, "u64": 8}
    off, align_max = 0, 1
    for name, ty in re.findall(r"pub (\w+): ([^,}]+)", body):
        ty = ty.strip()
        count = 1
        while ty.startswith("["):
            m = re.match(r"\[(.+); (\d+)\]$", ty)
            ty, count = m.group(1).strip(), count * int(m.group(2))
        size = 8 if ty.startswith("*") else sizes[ty]
        off = (off + size - 1) // size * size
        off += size * count
        align_max = max(align_max, size)
    return (off + align_max - 1) // align_max * align_max


def test_shim_mirrors_match_the_compiled_abi():
    shim = open(os.path.join(ROOT, "rust", "src", "gpu.rs")).read()
    L = pt.libptgpu()
    which = {"PtParams": 0, "PtCamera": 1, "PtTexture": 2, "PtMaterial": 3, "PtPerlin": 4, "PtSceneDesc": 5, "PtMotion": 9, "PtImage": 10, "PtOptions": 11}
    for name, idx in which.items():
        m = re.search(r"pub struct %s \{(.*?)\}" % name, shim, re.S)
        assert m, name
        assert _rust_struct_size(m.group(1)) == L.pt_abi_struct_size(idx), name
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ptgpu.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", header))
    imported = set(re.findall(r"fn (pt_[a-z0-9_]+)\(", shim))
    assert imported and imported <= declared, imported - declared
    assert {"pt_scene_create_multi", "pt_render", "pt_render_progressive", "pt_scene_destroy", "pt_last_error"} <= imported
    assert "pt_abi_version() }, %d)" % L.pt_abi_version() in shim
