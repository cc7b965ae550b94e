// restirpt_render — the headless stand-in of the reference executable (src/main.cpp:3-31: scene path, 1280x720,
// Renderer::exec): load a scene, run N frames of the chosen direct / indirect method through the host Renderer,
// print the frame rate the reference shows in its window title (src/Renderer.cpp:877-887) and write the last
// frame as a PNG (the reference's screenshot, src/Renderer.cpp:758-793).
//
// All device work goes through librestirpt.so; without a CUDA device the program stops with the library's error.
#include "Renderer.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

using namespace rpt;

namespace {

void usage() {
	std::fputs(
		"usage: restirpt_render [scene] [options]\n"
		"  scene                 a scene XML of the reference's dialect (default: res/model/VeachAjar/ajar.xml as the\n"
		"                        reference's main.cpp), or one of the built-in scenes: cornell, room, field\n"
		"  --size WxH            film size (default 1280x720, src/main.cpp:29)\n"
		"  --frames N            frames to render (default 64)\n"
		"  --direct M            none | naive | restir-di | visualize-as          (default none)\n"
		"  --indirect M          none | naive | restir-gi | restir-pt             (default restir-pt)\n"
		"  --shift S             reconnection | replay | hybrid                   (ReSTIR PT, default hybrid)\n"
		"  --temporal 0|1        temporal reuse (default 1)      --spatial 0|1   spatial reuse (default 1)\n"
		"  --cap N               reservoir M cap (default 20)    --rr-scale X    Russian-roulette scale (default 1)\n"
		"  --di-sample S         light | bsdf | both  (ReSTIR DI candidate sampling, default light)\n"
		"  --di-temporal 0|1     ReSTIR DI temporal reuse (default 0)   --di-spatial 0|1   spatial reuse (default 1)\n"
		"  --pipeline            RayTracing-pipeline mode of the naive direct pass (di_naive.rgen)\n"
		"  --accumulate          running mean over the frames (the reference's ground-truth mode)\n"
		"  --tonemap T           0 none | 1 filmic | 2 ACES (default 1)   --no-gamma\n"
		"  --seeds S             mt19937 (std::default_random_engine of the reference's platform, default) | hash2\n"
		"  --device D            CUDA device (default 0)\n"
		"  --out FILE.png        screenshot of the last frame (default: none)\n", stderr);
}

uint32_t hash2(uint32_t seed) {   // reference math.glsl:227-234
	seed = (seed ^ 61u) ^ (seed >> 16);
	seed *= 9u;
	seed = seed ^ (seed >> 4);
	seed *= 0x27d4eb2du;
	seed = seed ^ (seed >> 15);
	return seed;
}

int choice(const char* flag, const std::string& v, const std::vector<std::pair<const char*, int>>& table) {
	for (const auto& t : table) if (v == t.first) return t.second;
	std::string names;
	for (const auto& t : table) names += std::string(names.empty() ? "" : " | ") + t.first;
	throw std::runtime_error(std::string(flag) + ": '" + v + "' is not one of " + names);
}

} // namespace

int main(int argc, char** argv) {
	std::string scenePath = "res/model/VeachAjar/ajar.xml", out, seeds = "mt19937";
	uint32_t width = 1280, height = 720, frames = 64;
	int device = 0;
	RendererSettings settings;
	RptGRISSettings gris = { 2, 1.0f, 1, 1, 20 };
	RptDISettings di = { 0, 0, 0, 1 };   // src/TestReSTIR.h:29
	try {
		bool haveScene = false;
		for (int i = 1; i < argc; i++) {
			const std::string a = argv[i];
			auto value = [&]() -> std::string {
				if (i + 1 >= argc) throw std::runtime_error(a + " needs a value");
				return argv[++i];
			};
			if (a == "-h" || a == "--help") { usage(); return 0; }
			else if (a == "--size") {
				const std::string v = value();
				unsigned w = 0, h = 0;
				if (std::sscanf(v.c_str(), "%ux%u", &w, &h) != 2 || w == 0 || h == 0 || w > 16384 || h > 16384) throw std::runtime_error("--size: expected WxH");
				width = w; height = h;
			}
			else if (a == "--frames") { frames = uint32_t(std::max(1, std::atoi(value().c_str()))); }
			else if (a == "--direct") settings.directMethod = choice("--direct", value(), { { "none", RayTracingMethod::None }, { "naive", RayTracingMethod::Naive },
			                                                                                 { "restir-di", RayTracingMethod::ResampledDI }, { "visualize-as", RayTracingMethod::VisualizeAS } });
			else if (a == "--indirect") settings.indirectMethod = choice("--indirect", value(), { { "none", RayTracingMethod::None }, { "naive", RayTracingMethod::Naive },
			                                                                                       { "restir-gi", RayTracingMethod::ResampledGI }, { "restir-pt", RayTracingMethod::ResampledPT } });
			else if (a == "--shift") gris.shiftType = uint32_t(choice("--shift", value(), { { "reconnection", 0 }, { "replay", 1 }, { "hybrid", 2 } }));
			else if (a == "--temporal") gris.temporalReuse = std::atoi(value().c_str()) != 0;
			else if (a == "--spatial") gris.spatialReuse = std::atoi(value().c_str()) != 0;
			else if (a == "--cap") gris.cap = uint32_t(std::max(1, std::atoi(value().c_str())));
			else if (a == "--rr-scale") gris.rrScale = float(std::atof(value().c_str()));
			else if (a == "--di-sample") di.sampleType = uint32_t(choice("--di-sample", value(), { { "light", 0 }, { "bsdf", 1 }, { "both", 2 } }));
			else if (a == "--di-temporal") di.temporalReuse = std::atoi(value().c_str()) != 0;
			else if (a == "--di-spatial") di.spatialReuse = std::atoi(value().c_str()) != 0;
			else if (a == "--pipeline") settings.pipelineMode = 1;
			else if (a == "--accumulate") settings.accumulate = true;
			else if (a == "--tonemap") settings.toneMapping = choice("--tonemap", value(), { { "0", 0 }, { "1", 1 }, { "2", 2 } });
			else if (a == "--no-gamma") settings.correctGamma = false;
			else if (a == "--seeds") { seeds = value(); choice("--seeds", seeds, { { "mt19937", 0 }, { "hash2", 1 } }); }
			else if (a == "--device") device = std::atoi(value().c_str());
			else if (a == "--out") out = value();
			else if (!a.empty() && a[0] == '-') throw std::runtime_error("unknown option " + a);
			else if (!haveScene) { scenePath = a; haveScene = true; }
			else throw std::runtime_error("more than one scene given");
		}
	}
	catch (const std::exception& e) {
		std::fprintf(stderr, "restirpt_render: %s\n", e.what());
		usage();
		return 2;
	}

	try {
		Scene scene;
		if (scenePath == "cornell") makeCornellBox(scene);
		else if (scenePath == "room") makeAjarLikeRoom(scene, 380000, 1);
		else if (scenePath == "field") makeInstancedField(scene, 1, 3, 42);
		else scene.load(scenePath);
		scene.camera.setFilmSize(width, height);
		const RptSceneDesc d = scene.desc();
		std::fprintf(stderr, "scene %s: %u object triangles, %u light triangles, %u instances, %u materials, %u textures\n", scenePath.c_str(),
		             d.numIndices / 3, d.numTriangleLights, d.numInstances, d.numMaterials, d.numTextures);

		Renderer renderer(scene, width, height, device);
		renderer.settings = settings;
		renderer.grisSettings = gris;
		renderer.diSettings = di;
		std::vector<uint8_t> rgba(out.empty() ? 0 : size_t(width) * height * 4);
		std::mt19937 rng;
		const auto t0 = std::chrono::steady_clock::now();
		for (uint32_t f = 0; f < frames; f++) {
			const uint32_t seed = seeds == "hash2" ? hash2(f + 1) : uint32_t(rng());
			const bool last = f + 1 == frames;
			renderer.drawFrame(seed, (last && !out.empty()) ? rgba.data() : nullptr);
		}
		if (rpt_sync(renderer.frame()) != RPT_OK) throw std::runtime_error(std::string("rpt_sync: ") + rpt_last_error(renderer.ctx()));
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		std::printf("%u frames of %ux%u in %.3f s: %.1f frames/s\n", frames, width, height, sec, frames / sec);
		if (!out.empty()) {
			if (!writePNG(out, rgba.data(), width, height)) throw std::runtime_error("cannot write " + out);
			std::printf("wrote %s\n", out.c_str());
		}
	}
	catch (const std::exception& e) {
		std::fprintf(stderr, "restirpt_render: %s\n", e.what());
		return 1;
	}
	return 0;
}
