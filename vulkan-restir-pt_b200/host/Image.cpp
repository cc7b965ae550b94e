// Image decoding for the scene front-end: what the reference gets from stb_image through
// zvk::HostImage::createFromFile(path, Int8, filter, 4) (src/Resource.cpp:26) — any texture file as 8-bit RGBA.
//
// Written from the format specifications (ITU-T T.81 for JPEG, RFC 1950 / 1951 / 2083 for zlib / deflate / PNG):
//   JPEG  8-bit baseline, extended-sequential and progressive Huffman streams, 1 or 3 components, any sampling
//         factors, restart intervals.  Chroma planes at half resolution are interpolated with the usual triangle
//         filter (3:1 weights), YCbCr -> RGB in 16-bit fixed point.  Not handled: arithmetic coding, 12-bit, CMYK.
//   PNG   every colour type and bit depth, palette and tRNS transparency, Adam7 interlace; 16-bit samples keep their
//         high byte.
// The decoders are only the front-end's: both the CUDA library and the oracle receive the texels this file produces.
#include "Scene.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace rpt {

namespace {

struct DecodeError : std::runtime_error { using std::runtime_error::runtime_error; };
constexpr uint64_t MaxPixels = 1ull << 28;   // 16k x 16k: a damaged header must not turn into a 20 GB allocation

bool readFile(const std::string& p, std::vector<uint8_t>& out) {
	FILE* f = std::fopen(p.c_str(), "rb");
	if (!f) return false;
	std::fseek(f, 0, SEEK_END);
	const long n = std::ftell(f);
	std::fseek(f, 0, SEEK_SET);
	out.resize(n > 0 ? size_t(n) : 0);
	const bool ok = n >= 0 && std::fread(out.data(), 1, out.size(), f) == out.size();
	std::fclose(f);
	return ok;
}

// =============================================================================================================
// inflate (RFC 1951) inside a zlib wrapper (RFC 1950)
// =============================================================================================================
struct BitsLSB {
	const uint8_t* p; size_t n, pos = 0;
	uint32_t acc = 0; int cnt = 0;
	BitsLSB(const uint8_t* d, size_t len) : p(d), n(len) {}
	uint32_t get(int k) {
		while (cnt < k) {
			if (pos >= n) throw DecodeError("deflate stream ends early");
			acc |= uint32_t(p[pos++]) << cnt; cnt += 8;
		}
		const uint32_t v = k ? (acc & ((1u << k) - 1u)) : 0u;
		acc >>= k; cnt -= k;
		return v;
	}
	void alignByte() { acc = 0; cnt = 0; }
};

struct HuffLSB {   // canonical code, decoded bit by bit with the count / symbol tables of RFC 1951 3.2.2
	uint16_t count[16] = { 0 }, symbol[320] = { 0 };
	void build(const uint8_t* lengths, int n) {
		std::memset(count, 0, sizeof(count));
		for (int i = 0; i < n; i++) count[lengths[i]]++;
		count[0] = 0;
		uint16_t offs[16]; offs[1] = 0;
		for (int l = 1; l < 15; l++) offs[l + 1] = uint16_t(offs[l] + count[l]);
		for (int i = 0; i < n; i++) if (lengths[i]) symbol[offs[lengths[i]]++] = uint16_t(i);
	}
	int decode(BitsLSB& b) const {
		int code = 0, first = 0, index = 0;
		for (int l = 1; l <= 15; l++) {
			code |= int(b.get(1));
			const int c = count[l];
			if (code - c < first) return symbol[index + (code - first)];
			index += c; first += c; first <<= 1; code <<= 1;
		}
		throw DecodeError("bad deflate code");
	}
};

std::vector<uint8_t> inflateZlib(const uint8_t* d, size_t n, size_t expected) {
	if (n < 6 || (d[0] & 15) != 8 || ((d[0] << 8) | d[1]) % 31 != 0 || (d[1] & 0x20)) throw DecodeError("not a zlib stream");
	BitsLSB b(d + 2, n - 2);
	std::vector<uint8_t> out;
	out.reserve(expected);
	static const uint16_t lenBase[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
	static const uint8_t lenExtra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
	static const uint16_t distBase[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
	static const uint8_t distExtra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
	static const uint8_t clOrder[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
	for (bool last = false; !last;) {
		last = b.get(1) != 0;
		const uint32_t type = b.get(2);
		if (type == 0) {
			b.alignByte();
			if (b.pos + 4 > b.n) throw DecodeError("deflate stream ends early");
			const uint32_t len = b.p[b.pos] | (b.p[b.pos + 1] << 8), nlen = b.p[b.pos + 2] | (b.p[b.pos + 3] << 8);
			b.pos += 4;
			if ((len ^ 0xffffu) != nlen || b.pos + len > b.n) throw DecodeError("bad stored block");
			out.insert(out.end(), b.p + b.pos, b.p + b.pos + len);
			b.pos += len;
			continue;
		}
		if (type == 3) throw DecodeError("bad deflate block type");
		HuffLSB lit, dist;
		uint8_t lengths[320];
		if (type == 1) {
			for (int i = 0; i < 288; i++) lengths[i] = uint8_t(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
			lit.build(lengths, 288);
			for (int i = 0; i < 30; i++) lengths[i] = 5;
			dist.build(lengths, 30);
		}
		else {
			const int hlit = int(b.get(5)) + 257, hdist = int(b.get(5)) + 1, hclen = int(b.get(4)) + 4;
			uint8_t cl[19] = { 0 };
			for (int i = 0; i < hclen; i++) cl[clOrder[i]] = uint8_t(b.get(3));
			HuffLSB clh;
			clh.build(cl, 19);
			int i = 0;
			while (i < hlit + hdist) {
				const int sym = clh.decode(b);
				if (sym < 16) { lengths[i++] = uint8_t(sym); continue; }
				int rep; uint8_t val = 0;
				if (sym == 16) { if (i == 0) throw DecodeError("bad code lengths"); val = lengths[i - 1]; rep = 3 + int(b.get(2)); }
				else if (sym == 17) rep = 3 + int(b.get(3));
				else rep = 11 + int(b.get(7));
				if (i + rep > hlit + hdist) throw DecodeError("bad code lengths");
				while (rep--) lengths[i++] = val;
			}
			lit.build(lengths, hlit);
			dist.build(lengths + hlit, hdist);
		}
		for (;;) {
			const int sym = lit.decode(b);
			if (sym < 256) { out.push_back(uint8_t(sym)); continue; }
			if (sym == 256) break;
			if (sym > 285) throw DecodeError("bad length symbol");
			const size_t len = lenBase[sym - 257] + b.get(lenExtra[sym - 257]);
			const int ds = dist.decode(b);
			if (ds > 29) throw DecodeError("bad distance symbol");
			const size_t back = distBase[ds] + b.get(distExtra[ds]);
			if (back > out.size()) throw DecodeError("distance beyond the window");
			for (size_t k = 0, from = out.size() - back; k < len; k++) out.push_back(out[from + k]);
		}
	}
	return out;
}

// =============================================================================================================
// PNG
// =============================================================================================================
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

void unfilterRows(uint8_t* data, size_t rows, size_t rowBytes, size_t bpp) {   // rows of (filter byte + rowBytes), in place
	std::vector<uint8_t> zero(rowBytes, 0);
	const uint8_t* prev = zero.data();
	for (size_t y = 0; y < rows; y++) {
		uint8_t* row = data + y * (rowBytes + 1);
		const uint8_t ft = row[0];
		uint8_t* cur = row + 1;
		for (size_t i = 0; i < rowBytes; i++) {
			const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
			int pred = 0;
			switch (ft) {
			case 0: break;
			case 1: pred = a; break;
			case 2: pred = b; break;
			case 3: pred = (a + b) >> 1; break;
			case 4: {
				const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
				pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
				break;
			}
			default: throw DecodeError("bad PNG filter type");
			}
			cur[i] = uint8_t(cur[i] + pred);
		}
		prev = cur;
	}
}

void decodePNG(const std::vector<uint8_t>& file, HostImage& out) {
	size_t pos = 8;
	uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
	std::vector<uint8_t> idat, plte, trns;
	bool haveHdr = false, done = false;
	while (!done) {
		if (pos + 12 > file.size()) throw DecodeError("PNG ends early");
		const uint32_t len = be32(&file[pos]);
		const uint8_t* type = &file[pos + 4];
		const uint8_t* data = &file[pos + 8];
		if (pos + 12 + size_t(len) > file.size()) throw DecodeError("PNG chunk beyond the file");
		if (!std::memcmp(type, "IHDR", 4)) {
			if (len < 13) throw DecodeError("bad IHDR");
			w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
			if (data[10] != 0 || data[11] != 0 || interlace > 1 || w == 0 || h == 0 || w > 65535 || h > 65535 || uint64_t(w) * h > MaxPixels) throw DecodeError("unsupported PNG header");
			haveHdr = true;
		}
		else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
		else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
		else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
		else if (!std::memcmp(type, "IEND", 4)) done = true;
		pos += 12 + size_t(len);
	}
	if (!haveHdr) throw DecodeError("PNG without IHDR");
	const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	const bool depthOk = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
	                     (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
	                     ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
	if (!channels || !depthOk) throw DecodeError("unsupported PNG colour type / bit depth");
	if (ctype == 3 && plte.size() < 3) throw DecodeError("palette PNG without PLTE");
	const size_t bitsPerPixel = size_t(channels) * depth, bpp = std::max<size_t>(1, bitsPerPixel / 8);
	auto rowBytesOf = [&](size_t pw) { return (pw * bitsPerPixel + 7) / 8; };

	// the passes: one for a plain image, seven for Adam7 {x0, y0, dx, dy}
	static const int adam7[7][4] = { {0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2} };
	struct Pass { size_t x0, y0, dx, dy, pw, ph; };
	std::vector<Pass> passes;
	size_t total = 0;
	if (!interlace) passes.push_back({ 0, 0, 1, 1, w, h });
	else for (auto& a : adam7) {
		const size_t pw = (w > size_t(a[0])) ? (w - a[0] + a[2] - 1) / a[2] : 0, ph = (h > size_t(a[1])) ? (h - a[1] + a[3] - 1) / a[3] : 0;
		if (pw && ph) passes.push_back({ size_t(a[0]), size_t(a[1]), size_t(a[2]), size_t(a[3]), pw, ph });
	}
	for (auto& p : passes) total += p.ph * (rowBytesOf(p.pw) + 1);
	std::vector<uint8_t> raw = inflateZlib(idat.data(), idat.size(), total);
	if (raw.size() < total) throw DecodeError("PNG image data too short");

	out.width = w; out.height = h;
	out.rgba8.assign(size_t(w) * h * 4, 255);
	const int maxv = (1 << depth) - 1;
	auto sample = [&](const uint8_t* row, size_t idx) -> uint32_t {   // idx = sample index within the row
		if (depth == 8) return row[idx];
		if (depth == 16) return (uint32_t(row[2 * idx]) << 8) | row[2 * idx + 1];
		const size_t bit = idx * depth;
		return (row[bit >> 3] >> (8 - depth - (bit & 7))) & uint32_t(maxv);
	};
	auto to8 = [&](uint32_t v) -> uint8_t { return depth == 16 ? uint8_t(v >> 8) : depth == 8 ? uint8_t(v) : uint8_t(v * 255u / uint32_t(maxv)); };
	uint32_t keyG = 0x10000, keyR = 0x10000, keyGr = 0x10000, keyB = 0x10000;   // tRNS colour key of grey / RGB images
	if (ctype == 0 && trns.size() >= 2) keyG = (uint32_t(trns[0]) << 8) | trns[1];
	if (ctype == 2 && trns.size() >= 6) { keyR = (uint32_t(trns[0]) << 8) | trns[1]; keyGr = (uint32_t(trns[2]) << 8) | trns[3]; keyB = (uint32_t(trns[4]) << 8) | trns[5]; }

	size_t off = 0;
	for (auto& p : passes) {
		const size_t rb = rowBytesOf(p.pw);
		unfilterRows(raw.data() + off, p.ph, rb, bpp);
		for (size_t py = 0; py < p.ph; py++) {
			const uint8_t* row = raw.data() + off + py * (rb + 1) + 1;
			for (size_t px = 0; px < p.pw; px++) {
				uint8_t* o = &out.rgba8[((p.y0 + py * p.dy) * w + (p.x0 + px * p.dx)) * 4];
				switch (ctype) {
				case 0: { const uint32_t v = sample(row, px); o[0] = o[1] = o[2] = to8(v); o[3] = (v == keyG) ? 0 : 255; break; }
				case 2: {
					const uint32_t r = sample(row, 3 * px), g = sample(row, 3 * px + 1), b = sample(row, 3 * px + 2);
					o[0] = to8(r); o[1] = to8(g); o[2] = to8(b); o[3] = (r == keyR && g == keyGr && b == keyB) ? 0 : 255;
					break;
				}
				case 3: {
					const uint32_t i = sample(row, px);
					if (3 * size_t(i) + 2 >= plte.size()) throw DecodeError("palette index out of range");
					o[0] = plte[3 * i]; o[1] = plte[3 * i + 1]; o[2] = plte[3 * i + 2]; o[3] = i < trns.size() ? trns[i] : 255;
					break;
				}
				case 4: { o[0] = o[1] = o[2] = to8(sample(row, 2 * px)); o[3] = to8(sample(row, 2 * px + 1)); break; }
				default: { for (int c = 0; c < 4; c++) o[c] = to8(sample(row, 4 * px + c)); break; }
				}
			}
		}
		off += p.ph * (rb + 1);
	}
}

// =============================================================================================================
// JPEG (ITU-T T.81)
// =============================================================================================================
const uint8_t zigzag[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

struct JHuff {   // T.81 Annex F.2.2.3 tables plus an 9-bit direct look-up for the short codes
	bool defined = false;
	uint8_t vals[256];
	int mincode[17], maxcode[18], valptr[17];
	uint16_t fast[512];   // (length << 8) | symbol, 0 = longer than 9 bits
	void build(const uint8_t* counts, const uint8_t* symbols, int n) {
		std::memcpy(vals, symbols, size_t(n));
		std::memset(fast, 0, sizeof(fast));
		int code = 0, k = 0;
		for (int l = 1; l <= 16; l++) {
			valptr[l] = k; mincode[l] = code;
			for (int i = 0; i < counts[l - 1]; i++, k++, code++) {
				if (l <= 9) for (int fill = 0; fill < (1 << (9 - l)); fill++) fast[(code << (9 - l)) | fill] = uint16_t((l << 8) | symbols[k]);
			}
			maxcode[l] = counts[l - 1] ? code - 1 : -1;
			code <<= 1;
		}
		maxcode[17] = 0x7fffffff;
		defined = true;
	}
};

struct JBits {   // MSB-first reader over entropy-coded data: FF00 is a stuffed FF, any other marker ends the segment
	const uint8_t* p; size_t n, pos;
	uint32_t acc = 0; int cnt = 0;
	bool hitMarker = false;
	JBits(const uint8_t* d, size_t len, size_t start) : p(d), n(len), pos(start) {}
	void fill() {
		while (cnt <= 24) {
			uint32_t byte = 0;
			if (!hitMarker && pos < n) {
				byte = p[pos];
				if (byte == 0xff) {
					const uint8_t next = pos + 1 < n ? p[pos + 1] : 0xd9;
					if (next == 0) pos += 2;
					else { hitMarker = true; byte = 0; }
				}
				else pos++;
			}
			acc |= byte << (24 - cnt);
			cnt += 8;
		}
	}
	uint32_t peek(int k) { if (cnt < k) fill(); return acc >> (32 - k); }
	void skip(int k) { acc <<= k; cnt -= k; }
	uint32_t get(int k) { if (k == 0) return 0; const uint32_t v = peek(k); skip(k); return v; }
	void reset() { acc = 0; cnt = 0; hitMarker = false; }
};

int jdecode(JBits& b, const JHuff& h) {
	const uint32_t look = b.peek(16);
	const uint16_t f = h.fast[look >> 7];
	if (f) { b.skip(f >> 8); return f & 0xff; }
	for (int l = 10; l <= 16; l++) {
		const int code = int(look >> (16 - l));
		if (code <= h.maxcode[l]) { b.skip(l); return h.vals[h.valptr[l] + code - h.mincode[l]]; }
	}
	throw DecodeError("bad JPEG Huffman code");
}
int jextend(uint32_t v, int s) { return (s && v < (1u << (s - 1))) ? int(v) - (1 << s) + 1 : int(v); }   // T.81 F.2.2.1 EXTEND

struct JComp {
	int id = 0, h = 1, v = 1, tq = 0;
	int bw = 0, bh = 0;          // allocated blocks (whole MCUs)
	int cw = 0, ch = 0;          // the component's own size in samples
	std::vector<int16_t> coef;   // bw * bh * 64, natural order within a block
	std::vector<uint8_t> plane;  // (bw * 8) x (bh * 8) samples after the inverse transform
	int dcPred = 0;
};

struct JDecoder {
	const std::vector<uint8_t>& f;
	uint16_t qt[4][64] = { { 0 } };
	JHuff dc[4], ac[4];
	std::vector<JComp> comps;
	int width = 0, height = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0, restart = 0;
	bool progressive = false, haveFrame = false;
	explicit JDecoder(const std::vector<uint8_t>& file) : f(file) {}

	static int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

	void frame(const uint8_t* d, int len) {
		if (len < 6 || d[0] != 8) throw DecodeError("only 8-bit JPEG is supported");
		height = be16(d + 1); width = be16(d + 3);
		const int nc = d[5];
		if ((nc != 1 && nc != 3) || width == 0 || height == 0 || len < 6 + 3 * nc || uint64_t(width) * uint64_t(height) > MaxPixels) throw DecodeError("unsupported JPEG frame (components / size)");
		comps.resize(size_t(nc));
		for (int i = 0; i < nc; i++) {
			JComp& c = comps[size_t(i)];
			c.id = d[6 + 3 * i]; c.h = d[7 + 3 * i] >> 4; c.v = d[7 + 3 * i] & 15; c.tq = d[8 + 3 * i] & 3;
			if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4) throw DecodeError("bad JPEG sampling factors");
			hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v);
		}
		mcux = (width + 8 * hmax - 1) / (8 * hmax); mcuy = (height + 8 * vmax - 1) / (8 * vmax);
		for (JComp& c : comps) {
			c.bw = mcux * c.h; c.bh = mcuy * c.v;
			c.cw = (width * c.h + hmax - 1) / hmax; c.ch = (height * c.v + vmax - 1) / vmax;
			c.coef.assign(size_t(c.bw) * c.bh * 64, 0);
		}
		haveFrame = true;
	}

	// one 8x8 block of one scan
	void block(JBits& b, JComp& c, int16_t* q, const JHuff* hd, const JHuff* ha, int ss, int se, int ah, int al, int& eobrun) {
		if (!progressive) {
			const int t = jdecode(b, *hd);
			if (t > 15) throw DecodeError("bad JPEG DC size category");
			c.dcPred += jextend(b.get(t), t);
			q[0] = int16_t(c.dcPred);
			for (int k = 1; k < 64;) {
				const int rs = jdecode(b, *ha), r = rs >> 4, s = rs & 15;
				if (s == 0) { if (r != 15) break; k += 16; continue; }
				k += r;
				if (k > 63) throw DecodeError("JPEG coefficient index out of range");
				q[zigzag[k++]] = int16_t(jextend(b.get(s), s));
			}
			return;
		}
		if (ss == 0) {   // DC scans
			if (ah == 0) {
				const int t = jdecode(b, *hd);
				if (t > 15) throw DecodeError("bad JPEG DC size category");
				c.dcPred += jextend(b.get(t), t);
				q[0] = int16_t(c.dcPred * (1 << al));
			}
			else if (b.get(1)) q[0] = int16_t(q[0] | (1 << al));
			return;
		}
		if (ah == 0) {   // AC, first pass of a band (G.1.2.2)
			if (eobrun > 0) { eobrun--; return; }
			for (int k = ss; k <= se;) {
				const int rs = jdecode(b, *ha), r = rs >> 4, s = rs & 15;
				if (s == 0) {
					if (r < 15) { eobrun = (1 << r) - 1; if (r) eobrun += int(b.get(r)); break; }
					k += 16;
					continue;
				}
				k += r;
				if (k > 63) throw DecodeError("JPEG coefficient index out of range");
				q[zigzag[k++]] = int16_t(jextend(b.get(s), s) * (1 << al));
			}
			return;
		}
		// AC, refinement pass (G.1.2.3): one more bit for the coefficients that are already non-zero, new +-1 coefficients in between
		const int p1 = 1 << al, m1 = -(1 << al);
		auto refine = [&](int16_t& v) {
			if (b.get(1) && (v & p1) == 0) v = int16_t(v >= 0 ? v + p1 : v + m1);
		};
		int k = ss;
		if (eobrun == 0) {
			for (; k <= se; k++) {
				const int rs = jdecode(b, *ha);
				int r = rs >> 4, s = rs & 15;
				if (s) s = b.get(1) ? p1 : m1;
				else if (r != 15) { eobrun = 1 << r; if (r) eobrun += int(b.get(r)); break; }
				for (; k <= se; k++) {
					int16_t& v = q[zigzag[k]];
					if (v != 0) refine(v);
					else if (--r < 0) break;
				}
				if (s && k <= se) q[zigzag[k]] = int16_t(s);
			}
		}
		if (eobrun > 0) {
			for (; k <= se; k++) {
				int16_t& v = q[zigzag[k]];
				if (v != 0) refine(v);
			}
			eobrun--;
		}
	}

	size_t scan(size_t pos, int len) {   // pos = first byte of the SOS payload; returns the position after the entropy-coded data
		const uint8_t* d = &f[pos];
		const int ns = d[0];
		if (ns < 1 || ns > int(comps.size()) || len < 4 + 2 * ns) throw DecodeError("bad JPEG scan header");
		JComp* sc[4]; const JHuff* hd[4]; const JHuff* ha[4];
		for (int i = 0; i < ns; i++) {
			sc[i] = nullptr;
			for (JComp& c : comps) if (c.id == d[1 + 2 * i]) sc[i] = &c;
			if (!sc[i]) throw DecodeError("JPEG scan names an unknown component");
			hd[i] = &dc[(d[2 + 2 * i] >> 4) & 3]; ha[i] = &ac[d[2 + 2 * i] & 3];
		}
		const int ss = d[1 + 2 * ns], se = d[2 + 2 * ns], ah = d[3 + 2 * ns] >> 4, al = d[3 + 2 * ns] & 15;
		if (progressive ? (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13) : false) throw DecodeError("bad progressive scan parameters");
		for (int i = 0; i < ns; i++) {
			const bool needDc = !progressive || (ss == 0 && ah == 0), needAc = !progressive || ss > 0;
			if ((needDc && !hd[i]->defined) || (needAc && !ha[i]->defined)) throw DecodeError("JPEG scan uses an undefined Huffman table");
		}
		JBits b(f.data(), f.size(), pos + size_t(len));
		int eobrun = 0, rstLeft = restart, nextRst = 0;
		for (JComp& c : comps) c.dcPred = 0;
		auto restartCheck = [&]() {   // called before every MCU but the first
			if (restart == 0 || --rstLeft > 0) return;
			// to the marker: the rest of the current byte is padding
			b.reset();
			size_t p = b.pos;
			while (p + 1 < f.size() && !(f[p] == 0xff && f[p + 1] >= 0xd0 && f[p + 1] <= 0xd7)) {
				if (f[p] == 0xff && f[p + 1] != 0 && f[p + 1] != 0xff) throw DecodeError("JPEG restart marker missing");
				p++;
			}
			if (p + 1 >= f.size() || (f[p + 1] & 7) != nextRst) throw DecodeError("JPEG restart marker out of sequence");
			b.pos = p + 2;
			nextRst = (nextRst + 1) & 7;
			rstLeft = restart; eobrun = 0;
			for (JComp& c : comps) c.dcPred = 0;
		};
		if (ns == 1) {   // non-interleaved: the component's own blocks, row by row
			JComp& c = *sc[0];
			const int nbx = (c.cw + 7) / 8, nby = (c.ch + 7) / 8;
			bool first = true;
			for (int by = 0; by < nby; by++) for (int bx = 0; bx < nbx; bx++) {
				if (!first) restartCheck();
				first = false;
				block(b, c, &c.coef[(size_t(by) * c.bw + bx) * 64], hd[0], ha[0], ss, se, ah, al, eobrun);
			}
		}
		else {
			bool first = true;
			for (int my = 0; my < mcuy; my++) for (int mx = 0; mx < mcux; mx++) {
				if (!first) restartCheck();
				first = false;
				for (int i = 0; i < ns; i++) {
					JComp& c = *sc[i];
					for (int v = 0; v < c.v; v++) for (int h = 0; h < c.h; h++)
						block(b, c, &c.coef[(size_t(my * c.v + v) * c.bw + (mx * c.h + h)) * 64], hd[i], ha[i], ss, se, ah, al, eobrun);
				}
			}
		}
		// the next marker
		size_t p = b.pos;
		while (p + 1 < f.size() && !(f[p] == 0xff && f[p + 1] != 0 && f[p + 1] != 0xff && !(f[p + 1] >= 0xd0 && f[p + 1] <= 0xd7))) p++;
		return p;
	}

	void inverseTransform() {
		// separable inverse DCT in double precision (T.81 A.3.3), level shift, clamp
		double basis[8][8];
		for (int x = 0; x < 8; x++) for (int u = 0; u < 8; u++) basis[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0);
		for (JComp& c : comps) {
			const int pw = c.bw * 8;
			c.plane.assign(size_t(pw) * c.bh * 8, 0);
			const uint16_t* q = qt[c.tq];
			for (int by = 0; by < c.bh; by++) for (int bx = 0; bx < c.bw; bx++) {
				const int16_t* co = &c.coef[(size_t(by) * c.bw + bx) * 64];
				double in[64], tmp[64];
				bool acZero = true;
				for (int i = 0; i < 64; i++) { in[i] = double(co[i]) * q[i]; if (i && co[i]) acZero = false; }
				uint8_t* dst = &c.plane[size_t(by) * 8 * pw + size_t(bx) * 8];
				if (acZero) {
					const int v = std::min(255, std::max(0, int(std::floor(in[0] * 0.125 + 128.5))));
					for (int y = 0; y < 8; y++) std::memset(dst + size_t(y) * pw, v, 8);
					continue;
				}
				for (int v = 0; v < 8; v++) for (int x = 0; x < 8; x++) {   // rows
					double s = 0;
					for (int u = 0; u < 8; u++) s += basis[x][u] * in[v * 8 + u];
					tmp[v * 8 + x] = s;
				}
				for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {   // columns
					double s = 0;
					for (int v = 0; v < 8; v++) s += basis[y][v] * tmp[v * 8 + x];
					dst[size_t(y) * pw + x] = uint8_t(std::min(255, std::max(0, int(std::floor(s + 128.5)))));
				}
			}
		}
	}

	// component plane -> full resolution (width x height)
	std::vector<uint8_t> upsample(const JComp& c) const {
		std::vector<uint8_t> out(size_t(width) * height);
		const int pw = c.bw * 8;
		const int fx = hmax / c.h, fy = vmax / c.v;
		const bool exact = (hmax % c.h == 0) && (vmax % c.v == 0);
		if (exact && fx == 1 && fy == 1) {
			for (int y = 0; y < height; y++) std::memcpy(&out[size_t(y) * width], &c.plane[size_t(y) * pw], size_t(width));
			return out;
		}
		if (exact && fx == 2 && (fy == 1 || fy == 2) && c.cw > 2) {   // triangle filter: 3/4 of the nearer sample, 1/4 of the farther one, per axis
			// (planes of one or two columns are replicated below, as the common decoders do)
			std::vector<int> row(size_t(c.cw));
			for (int y = 0; y < height; y++) {
				const int sy = fy == 2 ? y >> 1 : y;
				const int ny = fy == 2 ? std::min(c.ch - 1, std::max(0, (y & 1) ? sy + 1 : sy - 1)) : sy;
				const uint8_t* a = &c.plane[size_t(std::min(sy, c.ch - 1)) * pw];
				const uint8_t* bb = &c.plane[size_t(ny) * pw];
				for (int x = 0; x < c.cw; x++) row[size_t(x)] = fy == 2 ? 3 * a[x] + bb[x] : 4 * a[x];   // x4
				uint8_t* o = &out[size_t(y) * width];
				for (int x = 0; x < width; x++) {
					const int sx = std::min(x >> 1, c.cw - 1);
					const int nx = std::min(c.cw - 1, std::max(0, (x & 1) ? sx + 1 : sx - 1));
					const int bias = fy == 2 ? ((x & 1) ? 7 : 8) : ((x & 1) ? 8 : 4);   // (the customary rounding pattern of this filter)
					o[x] = uint8_t((3 * row[size_t(sx)] + row[size_t(nx)] + bias) >> 4);
				}
			}
			return out;
		}
		for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) {   // any other ratio: nearest sample
			const int sx = std::min(c.cw - 1, x * c.h / hmax), sy = std::min(c.ch - 1, y * c.v / vmax);
			out[size_t(y) * width + x] = c.plane[size_t(sy) * pw + sx];
		}
		return out;
	}

	void decode(HostImage& out) {
		size_t pos = 2;
		bool done = false;
		while (!done) {
			while (pos < f.size() && f[pos] != 0xff) pos++;
			while (pos < f.size() && f[pos] == 0xff) pos++;
			if (pos >= f.size()) break;
			const uint8_t m = f[pos++];
			if (m == 0xd9) { done = true; break; }
			if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;
			if (pos + 2 > f.size()) throw DecodeError("JPEG ends early");
			const int len = be16(&f[pos]);
			if (len < 2 || pos + size_t(len) > f.size()) throw DecodeError("JPEG segment beyond the file");
			const uint8_t* d = &f[pos + 2];
			const int n = len - 2;
			switch (m) {
			case 0xc0: case 0xc1: case 0xc2:
				if (haveFrame) throw DecodeError("JPEG with more than one frame");
				progressive = m == 0xc2;
				frame(d, n);
				break;
			case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
				throw DecodeError("unsupported JPEG process (lossless / hierarchical / arithmetic)");
			case 0xc4:
				for (int i = 0; i + 17 <= n;) {
					const int tc = d[i] >> 4, th = d[i] & 15;
					int total = 0;
					for (int k = 0; k < 16; k++) total += d[i + 1 + k];
					if (tc > 1 || th > 3 || total > 256 || i + 17 + total > n) throw DecodeError("bad JPEG Huffman table");
					(tc ? ac : dc)[th].build(d + i + 1, d + i + 17, total);
					i += 17 + total;
				}
				break;
			case 0xdb:
				for (int i = 0; i < n;) {
					const int pq = d[i] >> 4, tq = d[i] & 15;
					if (tq > 3 || pq > 1 || i + 1 + 64 * (pq + 1) > n) throw DecodeError("bad JPEG quantisation table");
					for (int k = 0; k < 64; k++) qt[tq][zigzag[k]] = uint16_t(pq ? be16(d + i + 1 + 2 * k) : d[i + 1 + k]);
					i += 1 + 64 * (pq + 1);
				}
				break;
			case 0xdd:
				if (n >= 2) restart = be16(d);
				break;
			case 0xda:
				if (!haveFrame) throw DecodeError("JPEG scan before the frame header");
				pos = scan(pos + 2, n);
				continue;
			default: break;   // APPn, COM, ...
			}
			pos += size_t(len);
		}
		if (!haveFrame) throw DecodeError("JPEG without a frame");
		inverseTransform();
		out.width = uint32_t(width); out.height = uint32_t(height);
		out.rgba8.resize(size_t(width) * height * 4);
		if (comps.size() == 1) {
			const std::vector<uint8_t> y = upsample(comps[0]);
			for (size_t i = 0; i < y.size(); i++) { uint8_t* o = &out.rgba8[i * 4]; o[0] = o[1] = o[2] = y[i]; o[3] = 255; }
			return;
		}
		const std::vector<uint8_t> Y = upsample(comps[0]), Cb = upsample(comps[1]), Cr = upsample(comps[2]);
		auto clamp8 = [](int v) { return uint8_t(v < 0 ? 0 : v > 255 ? 255 : v); };
		for (size_t i = 0; i < Y.size(); i++) {   // JFIF YCbCr -> RGB, 16-bit fixed point
			const int y = Y[i], cb = int(Cb[i]) - 128, cr = int(Cr[i]) - 128;
			uint8_t* o = &out.rgba8[i * 4];
			o[0] = clamp8(y + ((91881 * cr + 32768) >> 16));
			o[1] = clamp8(y + ((-22554 * cb - 46802 * cr + 32768) >> 16));
			o[2] = clamp8(y + ((116130 * cb + 32768) >> 16));
			o[3] = 255;
		}
	}
};

} // namespace

bool readImage(const std::string& path, HostImage& out, std::string* error) {
	std::vector<uint8_t> file;
	if (!readFile(path, file)) { if (error) *error = "cannot read " + path; return false; }
	static const uint8_t pngSig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
	try {
		if (file.size() >= 8 && !std::memcmp(file.data(), pngSig, 8)) decodePNG(file, out);
		else if (file.size() >= 4 && file[0] == 0xff && file[1] == 0xd8) { JDecoder d(file); d.decode(out); }
		else if (file.size() >= 2 && file[0] == 'P' && file[1] == '6') return readPPM(path, out);
		else { if (error) *error = path + ": not a PNG, JPEG or binary PPM file"; return false; }
	}
	catch (const std::exception& e) {
		if (error) *error = path + ": " + e.what();
		return false;
	}
	return true;
}

} // namespace rpt
