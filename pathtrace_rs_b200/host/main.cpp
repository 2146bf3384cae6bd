// main.cpp — thin CLI shell with the reference's flags (src/main.rs:26-96), offline front-end only.
// The windowed front-end (glium_window.rs) needs a display and is out of scope; `-F frames`
// runs the same progressive loop headless (frame_num = 0..F-1 on one device-resident buffer, glium_window.rs:98-131).
// `-G` takes a device list: `-G 0`, `-G 0,2,5` or `-G 0-7` — the image is split over the listed GPUs inside the one
// `Scene::update` call.  Unknown flags and missing values are errors, as with clap in the reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "pathtrace.hpp"

static std::vector<int> parse_devices(const std::string& spec) {
    std::vector<int> out;
    size_t pos = 0;
    while (pos <= spec.size()) {
        const size_t comma = std::min(spec.find(',', pos), spec.size());
        const std::string item = spec.substr(pos, comma - pos);
        if (item.empty()) throw std::runtime_error("empty entry in device list '" + spec + "'");
        const size_t dash = item.find('-');
        char* end = nullptr;
        const long lo = std::strtol(item.c_str(), &end, 10);
        long hi = lo;
        if (dash != std::string::npos) {
            if (end != item.c_str() + dash) throw std::runtime_error("bad device range '" + item + "'");
            hi = std::strtol(item.c_str() + dash + 1, &end, 10);
        }
        if (*end != '\0' || lo < 0 || hi < lo || hi > 1023) throw std::runtime_error("bad device list entry '" + item + "'");
        for (long d = lo; d <= hi; ++d) out.push_back((int)d);
        pos = comma + 1;
    }
    return out;
}

int main(int argc, char** argv) {
    pathtrace::Params params;  // defaults: 1280x720, 4 spp, depth 10 (main.rs:78-85)
    std::string preset = "two_perlin_spheres";  // main.rs:87
    std::string output = "output.png";
    std::vector<int> devices{0};
    uint32_t frames = 1;
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            auto text = [&]() -> std::string {
                if (i + 1 >= argc) throw std::runtime_error("option '" + a + "' needs a value");
                return argv[++i];
            };
            auto number = [&]() -> uint32_t {
                const std::string v = text();
                char* end = nullptr;
                const unsigned long n = std::strtoul(v.c_str(), &end, 10);
                if (v.empty() || *end != '\0' || n > 0xffffffffUL) throw std::runtime_error("option '" + a + "': '" + v + "' is not a number");
                return (uint32_t)n;
            };
            if (a == "-W" || a == "--width") params.width = number();
            else if (a == "-H" || a == "--height") params.height = number();
            else if (a == "-S" || a == "--samples") params.samples = number();
            else if (a == "-D" || a == "--depth") params.max_depth = number();
            else if (a == "-R" || a == "--random") { params.random_seed = true; params.seed_salt = ((uint64_t)std::random_device{}() << 32) | std::random_device{}(); }
            else if (a == "-P" || a == "--preset") preset = text();
            else if (a == "-B" || a == "--bvh") params.use_bvh = true;
            else if (a == "-O" || a == "--offline") {}
            else if (a == "-o" || a == "--output") output = text();
            else if (a == "-G" || a == "--gpu") devices = parse_devices(text());
            else if (a == "-F" || a == "--frames") { frames = number(); if (frames == 0) throw std::runtime_error("--frames must be at least 1"); }
            else if (a == "-h" || a == "--help") {
                std::printf("Toy Path Tracer (B200)\n  -W/-H/-S/-D <n>  width/height/samples/depth\n  -R random seed  -P <preset>  -B bvh (rejected)  -O offline  -o <png>\n"
                            "  -G <devices>  GPU list: 0 | 0,2,5 | 0-7     -F <frames>  progressive accumulation, headless\n");
                return 0;
            } else {
                throw std::runtime_error("unexpected argument '" + a + "' (try --help)");
            }
        }
        pathtrace::offline::render_offline(preset, params, output, devices, frames);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());  // the reference panics via expect() / clap exits with usage
        return 101;
    }
    return 0;
}
