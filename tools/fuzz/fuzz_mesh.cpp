// fuzz driver: mutated PLY / STL files through readPLY / readSTL / triangulateRawMesh under ASan + UBSan
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include "MeshFormats.h"
using namespace rpt;
static std::string slurp(const char* p) { std::ifstream f(p, std::ios::binary); return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); }
int main(int argc, char** argv) {
	std::mt19937 rng(12345);
	int ok = 0, rejected = 0;
	for (int a = 1; a < argc; a++) {
		const std::string seed = slurp(argv[a]);
		const bool ply = std::string(argv[a]).find(".ply") != std::string::npos;
		for (int it = 0; it < 3000; it++) {
			std::string d = seed;
			const int edits = 1 + int(rng() % 6);
			for (int e = 0; e < edits && !d.empty(); e++) {
				const size_t pos = rng() % d.size();
				switch (rng() % 5) {
				case 0: d[pos] = char(rng()); break;
				case 1: d.erase(pos, 1 + rng() % 8); break;
				case 2: d.insert(pos, std::string(1 + rng() % 4, char('0' + rng() % 10))); break;
				case 3: d.resize(pos); break;
				default: d[pos] = "0123456789 -.\n"[rng() % 14]; break;
				}
			}
			const std::string tmp = "/tmp/fuzz/case." + std::string(ply ? "ply" : "stl");
			{ std::ofstream o(tmp, std::ios::binary); o.write(d.data(), std::streamsize(d.size())); }
			try {
				RawMesh m;
				if (ply) readPLY(tmp, m); else readSTL(tmp, m);
				std::vector<RptMeshVertex> v; std::vector<uint32_t> i;
				triangulateRawMesh(m, (it & 1) != 0, v, i);
				for (uint32_t k : i) if (k >= v.size()) { std::printf("index out of range\n"); return 2; }
				ok++;
			}
			catch (const std::exception&) { rejected++; }
		}
	}
	std::printf("fuzz: %d parsed, %d rejected, no crash\n", ok, rejected);
	return 0;
}
