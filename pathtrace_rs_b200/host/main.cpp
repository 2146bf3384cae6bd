// main.cpp — thin CLI shell with the reference's flags (src/main.rs:26-96), offline front-end only.
// The windowed front-end (glium_window.rs) needs a display and is out of scope; `-F frames`
// runs the same progressive loop headless (frame_num = 0..F-1 on one buffer, glium_window.rs:98-131).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "pathtrace.hpp"

int main(int argc, char** argv) {
    pathtrace::Params params;  // defaults: 1280x720, 4 spp, depth 10 (main.rs:78-85)
    std::string preset = "two_perlin_spheres";  // main.rs:87
    std::string output = "output.png";
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](uint32_t& dst) { if (i + 1 < argc) dst = (uint32_t)std::strtoul(argv[++i], nullptr, 10); };
        if (a == "-W" || a == "--width") value(params.width);
        else if (a == "-H" || a == "--height") value(params.height);
        else if (a == "-S" || a == "--samples") value(params.samples);
        else if (a == "-D" || a == "--depth") value(params.max_depth);
        else if (a == "-R" || a == "--random") { params.random_seed = true; params.seed_salt = ((uint64_t)std::random_device{}() << 32) | std::random_device{}(); }
        else if (a == "-P" || a == "--preset") { if (i + 1 < argc) preset = argv[++i]; }
        else if (a == "-B" || a == "--bvh") params.use_bvh = true;
        else if (a == "-O" || a == "--offline") {}
        else if (a == "-o" || a == "--output") { if (i + 1 < argc) output = argv[++i]; }
        else if (a == "-G" || a == "--gpu") { if (i + 1 < argc) device = std::atoi(argv[++i]); }
        else if (a == "-h" || a == "--help") {
            std::printf("Toy Path Tracer (B200)\n  -W/-H/-S/-D <n>  width/height/samples/depth\n  -R random seed  -P <preset>  -B bvh (rejected)  -O offline  -o <png>  -G <device>\n");
            return 0;
        }
    }
    try {
        pathtrace::offline::render_offline(preset, params, output, device);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());  // the reference panics via expect()
        return 101;
    }
    return 0;
}
