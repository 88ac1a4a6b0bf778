// benchmark — the reference's `benchmark` driver (reference src/benchmark.cpp:77-175, flags of src/args.cpp:7-50)
// for the CUDA backend: same flags, same scene, same summary lines, so scripts written against the reference's
// binary keep working with `-i cuda`.  The reference's own driver cannot be built offline (glm, OpenCL ICD and
// polyscope are fetched by CMake); INTEGRATION.md shows the `case Impl::CUDA` a maintainer adds there instead.
//
//   benchmark [-i cuda] [-d N[,N...]] [-n iter] [-w warmup] [-o dir] [-l] [-v] [--fp64]   reference flags (-d is a LIST of
//                                                          devices like the reference's, args.cpp:20-23; may be repeated)
//             [--gpus N]              shorthand for -d 0,1,...,N-1: the step runs slab-decomposed over N devices
//             [--scene 2cubes|dam] [--particles N] [--solver-iters I] [--surface on|off]     extensions
//             [--resident]            keep the particles on the device between frames (pbf_upload/step/download)
//             [--no-pin]              do not page-lock the particle vector (advance() then copies through pageable memory)
//
// Output directory (the reference's help text promises `cloud.ply, mesh.obj` but its save() only creates the
// directory, sph.hpp:188-196): cloud.ply = binary little-endian PLY of the final particles (x y z, r g b a, id),
// mesh.obj = the final marching-cubes mesh (v / vn, one face per vertex triple).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "pbf/sph.hpp"
#include "pbf/vec.hpp"
#include "pbf/cudasph.hpp"

using Clock = std::chrono::high_resolution_clock;
using Millis = std::chrono::duration<double, std::milli>;
using Particle = sph::Particle<size_t, float, pbf::vec>;
using Params = sph::SphParams<size_t, float, pbf::vec>;
using Result = sph::Result<size_t, float, pbf::vec>;

struct Options {
  std::string impl = "cuda", output = "./out_{impl}_{type}_{iter}", scene = "2cubes", saveState, loadState;
  std::vector<int> devices;  // -d/--devices (a list, args.cpp:20-23); empty = device 0
  size_t iterations = 200, warmup = 200, particles = 20000, solverIter = 6;
  bool list = false, verbose = false, fp64 = false, surface = true, resident = false, fountain = false, help = false, pin = true;
};

static void usage() {
  std::cout << "  benchmark {OPTIONS}\n\n    PBF sph benchmark\n\n  OPTIONS:\n"
               "      -h, --help              Display this help menu\n"
               "      -i[impl], --impl=[impl] Which implementation to use. One of: cuda  Default: cuda\n"
               "      -l, --list              List devices available for [impl] and exit\n"
               "      -v, --verbose           Show details such as device tree for [impl]\n"
               "      -d[dev...], --devices=[dev...] CUDA device ordinals (comma-separated and/or repeated); more than one =\n"
               "                              Z-curve slab decomposition over those GPUs. Default: 0\n"
               "      --gpus=[N]              Same as -d 0,1,...,N-1\n"
               "      -n[iter], --iter=[iter] How many iterations to run the simulation for. Default: 200\n"
               "      -w[warmup], --warmup=[warmup] Iterations to skip for warmup before timing starts. Default: 200\n"
               "      --fp64                  Use FP64 (not supported by the CUDA backend)\n"
               "      -o[out], --output=[out] Directory to write the final state (cloud.ply, mesh.obj) to; templates\n"
               "                              {iter}, {impl}, {type}. Default: ./out_{impl}_{type}_{iter}\n"
               "      --scene=[2cubes|dam]    Stock two-cube scene with the moving wall, or a dam-break block\n"
               "      --particles=[N]         Particle budget of the scene. Default: 20000\n"
               "      --save-state=[FILE]     Checkpoint: write the final particles (raw 56-byte sph::Particle records) to FILE\n"
               "      --load-state=[FILE]     Resume: start from the particles of a checkpoint instead of the scene's\n"
               "      --fountain              Pass a non-empty sph::Scene to every advance(): a well, a source, a drain and\n"
               "                              two queries (the scene of tests/helpers.py demo_scene)\n"
               "      --solver-iters=[I]      Solver iterations per step. Default: 6\n"
               "      --surface=[on|off]      Marching-cubes surface extraction each frame. Default: on\n"
               "      --resident              Keep particles on the device between frames\n"
               "      --no-pin                Do not page-lock the particle vector for advance()'s copies\n";
}

// "-n 5", "-n5", "--iter 5", "--iter=5"
static bool take(int argc, char **argv, int &i, const char *shortName, const char *longName, std::string &out) {
  const std::string a = argv[i];
  const std::string s = shortName ? std::string("-") + shortName : std::string(), l = std::string("--") + longName;
  if (!s.empty() && a.rfind(s, 0) == 0 && a.size() > s.size() && a[1] != '-') { out = a.substr(s.size()); return true; }
  if (a.rfind(l + "=", 0) == 0) { out = a.substr(l.size() + 1); return true; }
  if ((!s.empty() && a == s) || a == l) {
    if (i + 1 >= argc) throw std::runtime_error("missing value for " + a);
    out = argv[++i];
    return true;
  }
  return false;
}

static Options parse(int argc, char **argv) {
  Options o;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    std::string v;
    if (a == "-h" || a == "--help") o.help = true;
    else if (a == "-l" || a == "--list") o.list = true;
    else if (a == "-v" || a == "--verbose") o.verbose = true;
    else if (a == "--fp64") o.fp64 = true;
    else if (a == "--resident") o.resident = true;
    else if (a == "--no-pin") o.pin = false;
    else if (a == "--fountain") o.fountain = true;
    else if (take(argc, argv, i, "i", "impl", v)) o.impl = v;
    else if (take(argc, argv, i, "d", "devices", v)) {
      for (size_t b = 0; b <= v.size();) {  // "0,1,2" and repeated -d both extend the list
        const size_t e = std::min(v.find(',', b), v.size());
        if (e > b) o.devices.push_back(std::stoi(v.substr(b, e - b)));
        b = e + 1;
      }
    } else if (take(argc, argv, i, nullptr, "gpus", v)) {
      o.devices.clear();
      for (int d = 0; d < std::stoi(v); ++d) o.devices.push_back(d);
    }
    else if (take(argc, argv, i, "n", "iter", v)) o.iterations = std::stoull(v);
    else if (take(argc, argv, i, "w", "warmup", v)) o.warmup = std::stoull(v);
    else if (take(argc, argv, i, "o", "output", v)) o.output = v;
    else if (take(argc, argv, i, nullptr, "scene", v)) o.scene = v;
    else if (take(argc, argv, i, nullptr, "save-state", v)) o.saveState = v;
    else if (take(argc, argv, i, nullptr, "load-state", v)) o.loadState = v;
    else if (take(argc, argv, i, nullptr, "particles", v)) o.particles = std::stoull(v);
    else if (take(argc, argv, i, nullptr, "solver-iters", v)) o.solverIter = std::stoull(v);
    else if (take(argc, argv, i, nullptr, "surface", v)) o.surface = (v == "on" || v == "1" || v == "true");
    else throw std::runtime_error("Flag could not be matched: " + a);
  }
  return o;
}

static std::string replaceAll(std::string s, const std::string &from, const std::string &to) {
  for (size_t p = 0; (p = s.find(from, p)) != std::string::npos; p += to.size()) s.replace(p, from.size(), to);
  return s;
}

// Checkpoint / resume: "PBFSTATE", u64 count, then the particles exactly as they cross advance() (56-byte records).
static void saveState(const std::vector<Particle> &xs, const std::string &file) {
  std::ofstream out(file, std::ios::binary);
  const uint64_t n = xs.size();
  out.write("PBFSTATE", 8);
  out.write(reinterpret_cast<const char *>(&n), sizeof(n));
  out.write(reinterpret_cast<const char *>(xs.data()), std::streamsize(n * sizeof(Particle)));
  if (!out) throw std::runtime_error("cannot write " + file);
}
static std::vector<Particle> loadState(const std::string &file) {
  std::ifstream in(file, std::ios::binary);
  char magic[8];
  uint64_t n = 0;
  in.read(magic, 8);
  in.read(reinterpret_cast<char *>(&n), sizeof(n));
  if (!in || std::string(magic, 8) != "PBFSTATE") throw std::runtime_error(file + " is not a particle checkpoint");
  std::vector<Particle> xs(n);
  in.read(reinterpret_cast<char *>(xs.data()), std::streamsize(n * sizeof(Particle)));
  if (!in) throw std::runtime_error(file + " is truncated");
  return xs;
}

static void save(const Result &result, const std::vector<Particle> &xs, const std::string &dir) {
  if (dir.empty()) return;
  if (mkdir(dir.c_str(), 0755) != 0 && errno != EEXIST) throw std::runtime_error("cannot create " + dir);
  {
    std::ofstream ply(dir + "/cloud.ply", std::ios::binary);
    ply << "ply\nformat binary_little_endian 1.0\nelement vertex " << xs.size()
        << "\nproperty float x\nproperty float y\nproperty float z\nproperty float red\nproperty float green\n"
           "property float blue\nproperty float alpha\nproperty uint id\nend_header\n";
    for (const auto &p : xs) {
      const float rec[7] = {p.position.x, p.position.y, p.position.z, p.colour.x, p.colour.y, p.colour.z, p.colour.w};
      const uint32_t id = static_cast<uint32_t>(p.id);
      ply.write(reinterpret_cast<const char *>(rec), sizeof(rec));
      ply.write(reinterpret_cast<const char *>(&id), sizeof(id));
    }
  }
  {
    std::ofstream obj(dir + "/mesh.obj");
    obj << "# pbf-sph marching-cubes surface, " << result.mesh.vs.size() / 3 << " triangles\n";
    for (const auto &v : result.mesh.vs) obj << "v " << v.x << ' ' << v.y << ' ' << v.z << '\n';
    for (const auto &n : result.mesh.ns) obj << "vn " << n.x << ' ' << n.y << ' ' << n.z << '\n';
    for (size_t t = 0; t + 2 < result.mesh.vs.size(); t += 3)
      obj << "f " << t + 1 << "//" << t + 1 << ' ' << t + 2 << "//" << t + 2 << ' ' << t + 3 << "//" << t + 3 << '\n';
  }
}

int main(int argc, char *argv[]) {
  Options o;
  try {
    o = parse(argc, argv);
  } catch (const std::exception &e) {
    std::cerr << e.what() << std::endl;
    usage();
    return 1;
  }
  if (o.help) { usage(); return 0; }
  if (o.impl != "cuda") { std::cerr << "this build provides only the cuda implementation (got '" << o.impl << "')\n"; return 1; }
  if (o.fp64) { std::cerr << "FP64 not supported on CUDA" << std::endl; return 1; }  // like OCL, benchmark.cpp:140-141
  if (o.list) {
    pbf_ctx *probe = nullptr;
    for (int d = 0; pbf_create(&probe, 0.1f, d) == PBF_OK; ++d) {
      std::cout << "[" << d << "] CUDA device " << d << std::endl;
      pbf_destroy(probe);
    }
    return 0;
  }
  std::string output = replaceAll(replaceAll(replaceAll(o.output, "{iter}", std::to_string(o.iterations)), "{type}", "fp32"), "{impl}", o.impl);
  try {
    if (o.devices.empty()) o.devices.push_back(0);
    if (o.devices.size() > 1 && (o.resident || o.fountain))
      throw std::runtime_error("--resident and --fountain drive a single device");
    sph::cuda_impl::Solver<size_t, float, pbf::vec> solver(0.1f, o.devices, o.pin);  // h = 0.1, benchmark.cpp:160-163
    const float scaling = 500;  // benchmark.cpp:25
    auto [mc, param, particles] = o.scene == "dam"
        ? sph::damBreak<size_t, float, pbf::vec>(static_cast<size_t>(std::cbrt(double(o.particles)) + 0.5), o.solverIter, scaling)
        : sph::simpleConfigWith2Cubes<size_t, float, pbf::vec>(o.particles, o.solverIter, scaling);
    if (o.surface) param.surface = mc;  // marching cubes is ON in the stock benchmark (benchmark.cpp:29)
    if (!o.loadState.empty()) particles = loadState(o.loadState);
    const bool moving = o.scene != "dam";
    if (o.verbose) std::cout << "scene=" << o.scene << " particles=" << particles.size() << " solver-iters=" << o.solverIter
                             << " surface=" << (o.surface ? "on" : "off") << " resident=" << o.resident << std::endl;
    std::cout << "Using " << output << " for output" << std::endl;
    Result result;
    sph::Scene<size_t, float, pbf::vec> scene{};  // both reference drivers pass an empty Scene (benchmark.cpp:33,47)
    if (o.fountain) {
      using V3 = pbf::vec<3, float>;
      scene.wells.push_back({7, V3(300, 200, 300), 5000.0f});
      scene.sources.push_back({99, V3(500, 100, 500), V3(0, 1, 0), pbf::vec<4, float>(1, 0, 0, 1), 20.0f});
      scene.drains.push_back({1, V3(120, 20, 120), 40.0f, 1.0f});
      scene.queries.push_back({11, V3(150, 60, 150)});
      scene.queries.push_back({12, V3(900, 900, 900)});
      if (o.resident) throw std::runtime_error("--fountain goes through advance(config, scene, xs): not with --resident");
    }
    auto frameParams = [&](size_t frame) { return moving ? sph::applyMotionSinXCosZ(param, frame) : param; };
    auto advance = [&](size_t frame, const char *what) {
      try {
        if (!o.resident) { result = solver.advance(frameParams(frame), scene, particles); return; }
        const pbf_params p = decltype(solver)::toParams(frameParams(frame));
        if (pbf_step(solver.handle(), &p) != PBF_OK || pbf_sync(solver.handle()) != PBF_OK)
          throw std::runtime_error(pbf_last_error(solver.handle()));
      } catch (std::exception const &e) {
        std::cout << "Caught asynchronous exception at " << what << " frame" << frame << ":\n" << e.what() << "\n";
        throw;
      }
    };
    if (o.resident && pbf_upload(solver.handle(), reinterpret_cast<const pbf_particle *>(particles.data()), particles.size()) != PBF_OK)
      throw std::runtime_error(pbf_last_error(solver.handle()));
    for (size_t frame = 0; frame < o.warmup; ++frame) advance(frame, "warmup");
    std::vector<double> frameTime;
    const auto start = Clock::now();
    for (size_t frame = 0; frame < o.iterations; ++frame) {
      const auto t0 = Clock::now();
      advance(frame, "benchmark");
      frameTime.push_back(Millis(Clock::now() - t0).count());
    }
    const double seconds = Millis(Clock::now() - start).count() / 1000.0;
    if (o.resident) {
      uint64_t n = 0;
      if (pbf_download(solver.handle(), reinterpret_cast<pbf_particle *>(particles.data()), particles.size(), &n) != PBF_OK)
        throw std::runtime_error(pbf_last_error(solver.handle()));
      pbf_grid_info g;
      pbf_grid(solver.handle(), &g);
      const uint64_t nv = uint64_t(g.n_triangles) * 3;
      result.mesh.vs.resize(nv); result.mesh.ns.resize(nv); result.mesh.cs.resize(nv);
      if (nv) pbf_mesh_download(solver.handle(), reinterpret_cast<float *>(result.mesh.vs.data()), reinterpret_cast<float *>(result.mesh.ns.data()),
                                reinterpret_cast<float *>(result.mesh.cs.data()), nv);
    }
    const double frames = double(std::max<size_t>(1, frameTime.size()));
    const double mean = std::accumulate(frameTime.begin(), frameTime.end(), 0.0) / frames;
    double var = 0;
    for (double t : frameTime) var += (t - mean) * (t - mean);
    const auto [mn, mx] = std::minmax_element(frameTime.begin(), frameTime.end());
    // same labels as the reference's summary (benchmark.cpp:91-101)
    std::cout << "Benchmark completed after " << o.iterations << " frames:\n"
              << std::setprecision(4)  //
              << "Runtime              : " << seconds << " s\n"
              << "Framerate            : " << double(o.iterations) / seconds << " fps\n"
              << "Frame-time min       : " << (frameTime.empty() ? 0.0 : *mn) << " ms\n"
              << "Frame-time max       : " << (frameTime.empty() ? 0.0 : *mx) << " ms\n"
              << "Frame-time mean       : " << mean << " ms\n"
              << "Frame-time stdDev     : " << std::sqrt(var / frames) << " ms\n"
              << "Final Vertex count   : " << result.mesh.vs.size() << "\n"
              << "Final Particle count : " << particles.size() << " \n"
              << "Query answers        :" << [&] { std::string q; for (const auto &a : result.queries) q += " " + std::to_string(a.id) + ":" + std::to_string(a.neighbours.size()); return q.empty() ? std::string(" none") : q; }() << "\n"
              << "Particle-iterations/s: " << double(particles.size()) * double(o.solverIter) * double(o.iterations) / seconds << "\n"
              << std::endl;
    solver.unpin();  // the particle vector goes out of scope before the solver does: release its page-lock first
    if (!o.saveState.empty()) saveState(particles, o.saveState);
    save(result, particles, output);
    std::cout << "Results flushed." << std::endl;
  } catch (const std::exception &e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}
