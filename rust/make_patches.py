#!/usr/bin/env python
"""Regenerates rust/patches/*.diff: the reference-side change set that puts libptgpu.so behind `Scene::update`.

    python rust/make_patches.py [/root/reference]

Works on a scratch copy of the reference (nothing under /root/reference is touched), applies the edits below and writes one
unified diff per file (paths a/... b/..., `git apply -p1` / `patch -p1` from the reference's root).  The edits are all
additive: accessors for private fields, a `gpu_mask` field in Params with the `-G/--gpu` flag, and the GPU arm of
`Params::new_scene` / `Scene::update`.  tests/test_rust_patches.py applies the committed diffs to a fresh copy and checks
that they still apply cleanly and produce exactly these files.
"""
import difflib, os, shutil, sys, tempfile

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "patches")


def edit(path, pairs):
    s = open(path).read()
    for old, new in pairs:
        assert s.count(old) == 1, (path, old[:60], s.count(old))
        s = s.replace(old, new)
    open(path, "w").write(s)


EDITS = {
    "Cargo.toml": [
        ('edition = "2018"\n', 'edition = "2018"\nbuild = "build.rs"\n'),
        ('[features]\ncore_intrinsics = []\nbench = []\n',
         '[features]\ncore_intrinsics = []\nbench = []\n# the per-pixel path-tracing loop on NVIDIA B200 (sm_100a): build.rs compiles libptgpu.so with nvcc, src/gpu.rs binds it\ngpu = []\n'),
    ],
    "src/main.rs": [
        ('mod camera;\nmod collision;\nmod glium_window;\n', 'mod camera;\nmod collision;\nmod glium_window;\n#[cfg(feature = "gpu")]\nmod gpu;\n'),
        ('''            Arg::with_name("offline")
                .help("Don't create a preview render window")''',
         '''            Arg::with_name("gpu")
                .help("Render on these CUDA devices: 0 | 0,2,5 | 0-7 (needs the `gpu` feature)")
                .short("G")
                .long("gpu")
                .takes_value(true),
            Arg::with_name("offline")
                .help("Don't create a preview render window")'''),
        ('''        use_bvh: matches.is_present("bvh"),
    };
''', '''        use_bvh: matches.is_present("bvh"),
        gpu_mask: matches
            .value_of("gpu")
            .map_or(0, |spec| params::parse_gpu_mask(spec).expect("bad --gpu device list")),
    };
'''),
    ],
    "src/bench.rs": [
        ('    use_bvh: false,\n', '    use_bvh: false,\n    gpu_mask: 0,\n'),
    ],
    "src/params.rs": [
        ('''    pub random_seed: bool,
    pub use_bvh: bool,
}
''', '''    pub random_seed: bool,
    pub use_bvh: bool,
    /// bit i set = render on CUDA device i (all listed devices share every `Scene::update`, split by interleaved row tiles
    /// inside libptgpu); 0 = the CPU path.  A mask keeps `Params: Copy`.
    pub gpu_mask: u32,
}

/// `-G 0`, `-G 0,2,5`, `-G 0-7` -> device bit mask
pub fn parse_gpu_mask(spec: &str) -> Option<u32> {
    let mut mask = 0u32;
    for item in spec.split(',') {
        let mut ends = item.splitn(2, '-');
        let lo: u32 = ends.next()?.trim().parse().ok()?;
        let hi: u32 = match ends.next() {
            Some(hi) => hi.trim().parse().ok()?,
            None => lo,
        };
        if lo > hi || hi > 31 {
            return None;
        }
        for device in lo..=hi {
            mask |= 1 << device;
        }
    }
    Some(mask)
}
'''),
        ('''        Scene::new(hitable_list, sky)
    }
''', '''        #[allow(unused_mut)]
        let mut scene = Scene::new(hitable_list, sky);
        #[cfg(feature = "gpu")]
        {
            if self.gpu_mask != 0 {
                // the GPU arm: flatten the sphere list once, upload it to every listed device (panics on use_bvh or on a
                // hitable that is not a Sphere / MovingSphere, like SpheresSoA::new); Scene::update then forwards to it
                let devices: Vec<i32> = (0..32).filter(|d| self.gpu_mask & (1 << d) != 0).collect();
                scene.attach_gpu(&storage.perlin_noise, &devices);
            }
        }
        #[cfg(not(feature = "gpu"))]
        assert_eq!(self.gpu_mask, 0, "--gpu needs a build with `--features gpu`");
        scene
    }
'''),
    ],
    "src/camera.rs": [
        ('''    pub fn get_ray<T: Rng>(&self, s: f32, t: f32, rng: &mut T) -> Ray {''',
         '''    /// the ten private fields in the order of libptgpu's `PtCamera` (include/ptgpu.h)
    #[cfg(feature = "gpu")]
    pub(crate) fn to_ffi(&self) -> crate::gpu::PtCamera {
        let a = |v: Vec3| [v.x, v.y, v.z];
        crate::gpu::PtCamera {
            origin: a(self.origin),
            lower_left_corner: a(self.lower_left_corner),
            horizontal: a(self.horizontal),
            vertical: a(self.vertical),
            u: a(self.u),
            v: a(self.v),
            w: a(self.w),
            time0: self.time0,
            time1: self.time1,
            lens_radius: self.lens_radius,
        }
    }

    pub fn get_ray<T: Rng>(&self, s: f32, t: f32, rng: &mut T) -> Ray {'''),
    ],
    "src/scene.rs": [
        ('''    sky: Option<Vec3>,
    ray_count: AtomicUsize,
}
''', '''    sky: Option<Vec3>,
    ray_count: AtomicUsize,
    /// device copies of `world` (feature `gpu`, `Params::gpu_mask != 0`): `update` forwards to them
    #[cfg(feature = "gpu")]
    gpu: Option<crate::gpu::GpuScene>,
}
'''),
        ('''            sky,
            ray_count: AtomicUsize::new(0),
        }
    }
''', '''            sky,
            ray_count: AtomicUsize::new(0),
            #[cfg(feature = "gpu")]
            gpu: None,
        }
    }

    #[cfg(feature = "gpu")]
    pub(crate) fn world(&self) -> &Hitable<'a> {
        &self.world
    }

    /// `Scene.sky` (the constant sky colour, if any); `sky(&self, &Ray)` below is the lookup
    #[cfg(feature = "gpu")]
    pub(crate) fn sky_colour(&self) -> Option<Vec3> {
        self.sky
    }

    #[cfg(feature = "gpu")]
    pub(crate) fn attach_gpu(&mut self, perlin: &crate::perlin::Perlin, devices: &[i32]) {
        self.gpu = Some(crate::gpu::GpuScene::new(self, perlin, devices));
    }
'''),
    ],
    "src/collision/hitable_list.rs": [
        ('''    pub fn new(hitables: Vec<Hitable>) -> HitableList {
        HitableList { hitables }
    }
''', '''    pub fn new(hitables: Vec<Hitable>) -> HitableList {
        HitableList { hitables }
    }

    pub(crate) fn hitables(&self) -> &[Hitable<'a>] {
        &self.hitables
    }
'''),
    ],
    "src/perlin.rs": [
        ('''    #[inline]
    fn interpolate(c: &[[[Vec3; 2]; 2]; 2], u: f32, v: f32, w: f32) -> f32 {''',
         '''    /// (randvec, perm_x, perm_y, perm_z): 256 entries each
    pub(crate) fn tables(&self) -> (&[Vec3], &[u32], &[u32], &[u32]) {
        (&self.randvec, &self.perm_x, &self.perm_y, &self.perm_z)
    }

    #[inline]
    fn interpolate(c: &[[[Vec3; 2]; 2]; 2], u: f32, v: f32, w: f32) -> f32 {'''),
    ],
    "src/texture.rs": [
        ('''    pub fn value(&self, u: f32, v: f32) -> Vec3 {
        let i = (u * self.width as f32) as i32;''',
         '''    /// (width, height, packed 8-bit RGB rows, row 0 = top)
    pub(crate) fn raw(&self) -> (u32, u32, &[u8]) {
        (self.width, self.height, &self.data)
    }

    pub fn value(&self, u: f32, v: f32) -> Vec3 {
        let i = (u * self.width as f32) as i32;'''),
    ],
    "src/collision/moving_sphere.rs": [
        ('''    #[inline]
    pub fn radius(&self) -> f32 {
        self.radius
    }
''', '''    #[inline]
    pub fn radius(&self) -> f32 {
        self.radius
    }

    /// the constructor's arguments back: (centre0, centre1, time0, time1)
    pub(crate) fn endpoints(&self) -> (Vec3, Vec3, f32, f32) {
        (
            self.centre_start,
            self.centre_start + self.centre_delta,
            self.time_start,
            self.time_start + 1.0 / self.inv_time_delta,
        )
    }
'''),
    ],
}

# Scene::update: the forwarding arm goes at the top of the function body
SCENE_UPDATE_OLD = '''        buffer: &mut [(f32, f32, f32)],
    ) -> usize {
'''
SCENE_UPDATE_NEW = '''        buffer: &mut [(f32, f32, f32)],
    ) -> usize {
        #[cfg(feature = "gpu")]
        {
            if let Some(gpu) = &self.gpu {
                return gpu.update(params, camera, frame_num, buffer); // -> pt_render: same arguments, same return
            }
        }
'''


def main():
    tmp = tempfile.mkdtemp(prefix="ref_patch_")
    a, b = os.path.join(tmp, "a"), os.path.join(tmp, "b")
    for d in (a, b):
        os.makedirs(d)
        shutil.copy(os.path.join(REF, "Cargo.toml"), d)
        shutil.copytree(os.path.join(REF, "src"), os.path.join(d, "src"))
    for rel, pairs in EDITS.items():
        edit(os.path.join(b, rel), pairs)
    edit(os.path.join(b, "src/scene.rs"), [(SCENE_UPDATE_OLD, SCENE_UPDATE_NEW)])
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        if f.endswith(".diff"):
            os.remove(os.path.join(OUT, f))
    for n, rel in enumerate(sorted(EDITS), 1):
        old = open(os.path.join(a, rel)).read().splitlines(keepends=True)
        new = open(os.path.join(b, rel)).read().splitlines(keepends=True)
        diff = "".join(difflib.unified_diff(old, new, "a/" + rel, "b/" + rel))
        assert diff, rel
        name = "%02d-%s.diff" % (n, rel.replace("/", "-").replace(".", "_"))
        open(os.path.join(OUT, name), "w").write(diff)
        print("wrote", name, "(+%d lines)" % sum(1 for l in diff.splitlines() if l.startswith("+") and not l.startswith("+++")))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
