// io.cpp -- file formats of the reference's examples and fixtures (include/OptB200IO.h): .imagedump, PNG, PLY/OBJ/OFF meshes,
// marker / constraint text files, the shape_from_shading parameter dump.  Plain C++ + zlib (the reference links -lz too,
// examples/shared/make_template.inc:15); no CUDA in this file.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <zlib.h>
#include "../../include/OptB200IO.h"

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

bool readAll(const char* path, std::vector<unsigned char>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == buf.size();
}
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void putBe32(std::vector<unsigned char>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
std::string lowerExt(const char* path) {
    std::string s(path); size_t d = s.find_last_of('.');
    std::string e = d == std::string::npos ? "" : s.substr(d + 1);
    for (auto& c : e) c = (char)tolower(c);
    return e;
}
int paeth(int a, int b, int c) { int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
}  // namespace

extern "C" {

const char* OptIO_LastError(void) { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------------ .imagedump
int OptIO_ReadImageDump(const char* path, int clampInfinity, OptIO_Image* out) {
    if (!path || !out) return fail(-3, "null argument");
    std::vector<unsigned char> buf;
    if (!readAll(path, buf)) return fail(-1, std::string("cannot read ") + path);
    if (buf.size() < 16) return fail(-2, "imagedump: shorter than its header");
    int32_t hdr[4]; memcpy(hdr, buf.data(), 16);
    const int w = hdr[0], h = hdr[1], c = hdr[2], t = hdr[3];
    if (w <= 0 || h <= 0 || c <= 0 || (t != 0 && t != 1)) return fail(-2, "imagedump: bad header");
    const size_t es = t == 0 ? 4 : 1, bytes = es * (size_t)c * (size_t)w * (size_t)h;
    if (buf.size() < 16 + bytes) return fail(-2, "imagedump: truncated data");
    void* data = malloc(bytes);
    memcpy(data, buf.data() + 16, bytes);
    if (t == 0 && clampInfinity) {   // SimpleBuffer.cpp:29-40 (the loader clamps the first width*height floats)
        float* f = (float*)data;
        for (size_t i = 0; i < (size_t)w * h; ++i) if (std::isinf(f[i])) f[i] = f[i] > 0 ? FLT_MAX : -10000.0f;
    }
    out->width = w; out->height = h; out->channels = c; out->type = t; out->data = data;
    return 0;
}
int OptIO_WriteImageDump(const char* path, const OptIO_Image* img) {
    if (!path || !img || !img->data) return fail(-3, "null argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(-1, std::string("cannot write ") + path);
    int32_t hdr[4] = { img->width, img->height, img->channels, img->type };
    const size_t es = img->type == 0 ? 4 : 1, bytes = es * (size_t)img->channels * (size_t)img->width * (size_t)img->height;
    bool ok = fwrite(hdr, 4, 4, f) == 4 && fwrite(img->data, 1, bytes, f) == bytes;
    fclose(f);
    return ok ? 0 : fail(-1, "short write");
}
void OptIO_FreeImage(OptIO_Image* img) { if (img && img->data) { free(img->data); img->data = nullptr; } }

int OptIO_ReadSFSParameters(const char* path, float out40[40]) {
    if (!path || !out40) return fail(-3, "null argument");
    std::vector<unsigned char> buf;
    if (!readAll(path, buf)) return fail(-1, std::string("cannot read ") + path);
    if (buf.size() < 156) return fail(-2, "SFSSolverParameters: expected at least 156 bytes (36 floats + 3 uints)");
    memcpy(out40, buf.data(), 36 * 4);
    uint32_t u[3]; memcpy(u, buf.data() + 144, 12);
    for (int i = 0; i < 3; ++i) out40[36 + i] = (float)u[i];
    out40[39] = 0.f;
    return 0;
}

// ------------------------------------------------------------------------------------------------ PNG
int OptIO_ReadPNG(const char* path, OptIO_Image* out) {
    if (!path || !out) return fail(-3, "null argument");
    std::vector<unsigned char> buf;
    if (!readAll(path, buf)) return fail(-1, std::string("cannot read ") + path);
    static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    if (buf.size() < 8 || memcmp(buf.data(), sig, 8)) return fail(-2, "not a PNG file");
    uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, palette, trns;
    size_t pos = 8; bool haveHdr = false;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]); const char* type = (const char*)&buf[pos + 4];
        if (pos + 12 + len > buf.size()) return fail(-2, "PNG: truncated chunk");
        const unsigned char* d = &buf[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) return fail(-2, "PNG: bad IHDR");
            w = be32(d); h = be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; haveHdr = true;
        } else if (!memcmp(type, "PLTE", 4)) palette.assign(d, d + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(d, d + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + len;
    }
    if (!haveHdr || w == 0 || h == 0) return fail(-2, "PNG: no IHDR");
    if (interlace) return fail(-2, "PNG: interlaced images are not supported");
    int ch = 0;
    switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 3: ch = 1; break; case 4: ch = 2; break; case 6: ch = 4; break; default: return fail(-2, "PNG: bad colour type"); }
    if (!(depth == 8 || depth == 16 || (ctype == 3 && (depth == 1 || depth == 2 || depth == 4)) || (ctype == 0 && (depth == 1 || depth == 2 || depth == 4))))
        return fail(-2, "PNG: unsupported bit depth");
    const size_t bpp = std::max<size_t>(1, (size_t)ch * depth / 8);           // bytes per complete pixel for the filters
    const size_t rowBytes = ((size_t)w * ch * depth + 7) / 8;
    std::vector<unsigned char> raw((rowBytes + 1) * (size_t)h);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return fail(-2, "PNG: zlib stream is corrupt");
    std::vector<unsigned char> prev(rowBytes, 0), cur(rowBytes);
    unsigned char* rgba = (unsigned char*)malloc((size_t)w * h * 4);
    for (uint32_t y = 0; y < h; ++y) {
        const unsigned char* src = &raw[(rowBytes + 1) * y];
        const int ft = src[0];
        for (size_t i = 0; i < rowBytes; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = src[1 + i];
            switch (ft) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) / 2; break; case 4: v += paeth(a, b, c); break;
                          default: free(rgba); return fail(-2, "PNG: bad filter type"); }
            cur[i] = (unsigned char)v;
        }
        for (uint32_t x = 0; x < w; ++x) {
            unsigned char px[4] = { 0, 0, 0, 255 };
            auto sample = [&](int k) -> int {   // k-th sample of pixel x, scaled to 8 bits
                if (depth == 8) return cur[(size_t)x * ch + k];
                if (depth == 16) return cur[((size_t)x * ch + k) * 2];
                const size_t bit = ((size_t)x * ch + k) * depth; const int v = (cur[bit / 8] >> (8 - depth - (bit % 8))) & ((1 << depth) - 1);
                return ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
            };
            if (ctype == 0) { px[0] = px[1] = px[2] = (unsigned char)sample(0); }
            else if (ctype == 2) { px[0] = (unsigned char)sample(0); px[1] = (unsigned char)sample(1); px[2] = (unsigned char)sample(2); }
            else if (ctype == 3) {
                const int idx = sample(0);
                if ((size_t)idx * 3 + 2 < palette.size()) { px[0] = palette[idx * 3]; px[1] = palette[idx * 3 + 1]; px[2] = palette[idx * 3 + 2]; }
                if ((size_t)idx < trns.size()) px[3] = trns[idx];
            }
            else if (ctype == 4) { px[0] = px[1] = px[2] = (unsigned char)sample(0); px[3] = (unsigned char)sample(1); }
            else { px[0] = (unsigned char)sample(0); px[1] = (unsigned char)sample(1); px[2] = (unsigned char)sample(2); px[3] = (unsigned char)sample(3); }
            memcpy(&rgba[((size_t)y * w + x) * 4], px, 4);
        }
        std::swap(prev, cur);
    }
    out->width = (int)w; out->height = (int)h; out->channels = 4; out->type = 1; out->data = rgba;
    return 0;
}
int OptIO_WritePNG(const char* path, const OptIO_Image* img) {
    if (!path || !img || !img->data || img->type != 1 || !(img->channels == 1 || img->channels == 3 || img->channels == 4)) return fail(-3, "WritePNG: uint8 image with 1, 3 or 4 channels expected");
    const size_t rowBytes = (size_t)img->width * img->channels;
    std::vector<unsigned char> raw((rowBytes + 1) * (size_t)img->height);
    for (int y = 0; y < img->height; ++y) { raw[(rowBytes + 1) * y] = 0; memcpy(&raw[(rowBytes + 1) * y + 1], (const unsigned char*)img->data + rowBytes * y, rowBytes); }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<unsigned char> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return fail(-2, "WritePNG: deflate failed");
    comp.resize(clen);
    std::vector<unsigned char> file = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    auto chunk = [&](const char* type, const std::vector<unsigned char>& d) {
        putBe32(file, (uint32_t)d.size());
        const size_t at = file.size();
        file.insert(file.end(), type, type + 4); file.insert(file.end(), d.begin(), d.end());
        putBe32(file, (uint32_t)crc32(0L, &file[at], (uInt)(4 + d.size())));
    };
    std::vector<unsigned char> ih; putBe32(ih, (uint32_t)img->width); putBe32(ih, (uint32_t)img->height);
    ih.push_back(8); ih.push_back(img->channels == 1 ? 0 : (img->channels == 3 ? 2 : 6)); ih.push_back(0); ih.push_back(0); ih.push_back(0);
    chunk("IHDR", ih); chunk("IDAT", comp); chunk("IEND", {});
    FILE* f = fopen(path, "wb");
    if (!f) return fail(-1, std::string("cannot write ") + path);
    const bool ok = fwrite(file.data(), 1, file.size(), f) == file.size();
    fclose(f);
    return ok ? 0 : fail(-1, "short write");
}

// ------------------------------------------------------------------------------------------------ meshes
static int finishMesh(const std::vector<float>& v, const std::vector<int>& fs, const std::vector<int>& fi, OptIO_Mesh* out) {
    const int nv = (int)(v.size() / 3);
    for (int i : fi) if (i < 0 || i >= nv) return fail(-2, "mesh: face index out of range");
    out->nVertices = nv; out->nFaces = (int)fs.size() - 1;
    out->vertices = (float*)malloc(sizeof(float) * std::max<size_t>(1, v.size())); memcpy(out->vertices, v.data(), sizeof(float) * v.size());
    out->faceStart = (int*)malloc(sizeof(int) * fs.size()); memcpy(out->faceStart, fs.data(), sizeof(int) * fs.size());
    out->faceIndices = (int*)malloc(sizeof(int) * std::max<size_t>(1, fi.size())); memcpy(out->faceIndices, fi.data(), sizeof(int) * fi.size());
    return 0;
}
static size_t plyTypeSize(const std::string& t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
static double plyScalar(const unsigned char* p, const std::string& t) {
    if (t == "char" || t == "int8") return (double)*(const int8_t*)p;
    if (t == "uchar" || t == "uint8") return (double)*p;
    if (t == "short" || t == "int16") { int16_t v; memcpy(&v, p, 2); return v; }
    if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, p, 2); return v; }
    if (t == "int" || t == "int32") { int32_t v; memcpy(&v, p, 4); return v; }
    if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, p, 4); return v; }
    if (t == "float" || t == "float32") { float v; memcpy(&v, p, 4); return v; }
    double v; memcpy(&v, p, 8); return v;
}
static int readPLY(const std::vector<unsigned char>& buf, OptIO_Mesh* out) {
    struct Prop { std::string name, type, countType; bool list; };
    struct Elem { std::string name; long count; std::vector<Prop> props; };
    std::vector<Elem> elems; bool binary = false;
    size_t pos = 0; bool ended = false; bool first = true;
    while (pos < buf.size() && !ended) {
        size_t e = pos; while (e < buf.size() && buf[e] != '\n') ++e;
        std::string line((const char*)&buf[pos], e - pos); pos = e + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line); std::string w; ss >> w;
        if (first) { if (w != "ply") return fail(-2, "PLY: missing magic"); first = false; continue; }
        if (w == "format") { std::string f; ss >> f; if (f == "binary_little_endian") binary = true; else if (f != "ascii") return fail(-2, "PLY: only ascii and binary_little_endian are supported"); }
        else if (w == "element") { Elem el; ss >> el.name >> el.count; elems.push_back(el); }
        else if (w == "property") {
            if (elems.empty()) return fail(-2, "PLY: property before element");
            Prop p; std::string t; ss >> t;
            if (t == "list") { p.list = true; ss >> p.countType >> p.type >> p.name; } else { p.list = false; p.type = t; ss >> p.name; }
            if (!plyTypeSize(p.type) || (p.list && !plyTypeSize(p.countType))) return fail(-2, "PLY: unknown property type");
            elems.back().props.push_back(p);
        } else if (w == "end_header") ended = true;
    }
    if (!ended) return fail(-2, "PLY: no end_header");
    std::vector<float> v; std::vector<int> fs{ 0 }, fi;
    std::istringstream text;
    if (!binary) text.str(std::string((const char*)&buf[pos], buf.size() - pos));
    for (auto& el : elems) {
        for (long i = 0; i < el.count; ++i) {
            float xyz[3] = { 0, 0, 0 };
            for (auto& p : el.props) {
                long n = 1;
                if (p.list) {
                    if (binary) { const size_t s = plyTypeSize(p.countType); if (pos + s > buf.size()) return fail(-2, "PLY: truncated"); n = (long)plyScalar(&buf[pos], p.countType); pos += s; }
                    else { double d; if (!(text >> d)) return fail(-2, "PLY: truncated"); n = (long)d; }
                }
                for (long k = 0; k < n; ++k) {
                    double d;
                    if (binary) { const size_t s = plyTypeSize(p.type); if (pos + s > buf.size()) return fail(-2, "PLY: truncated"); d = plyScalar(&buf[pos], p.type); pos += s; }
                    else if (!(text >> d)) return fail(-2, "PLY: truncated");
                    if (el.name == "vertex" && !p.list) { if (p.name == "x") xyz[0] = (float)d; else if (p.name == "y") xyz[1] = (float)d; else if (p.name == "z") xyz[2] = (float)d; }
                    if (el.name == "face" && p.list && (p.name == "vertex_indices" || p.name == "vertex_index")) fi.push_back((int)d);
                }
            }
            if (el.name == "vertex") v.insert(v.end(), xyz, xyz + 3);
            if (el.name == "face") fs.push_back((int)fi.size());
        }
    }
    return finishMesh(v, fs, fi, out);
}
static int readOBJ(const std::vector<unsigned char>& buf, OptIO_Mesh* out) {
    std::istringstream in(std::string((const char*)buf.data(), buf.size()));
    std::vector<float> v; std::vector<int> fs{ 0 }, fi; std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line); std::string w; ss >> w;
        if (w == "v") { float x, y, z; if (ss >> x >> y >> z) { v.push_back(x); v.push_back(y); v.push_back(z); } }
        else if (w == "f") {
            std::string tok;
            while (ss >> tok) {
                int idx = atoi(tok.c_str());                       // "i", "i/t", "i//n", "i/t/n"
                if (idx < 0) idx = (int)(v.size() / 3) + idx + 1;  // relative index
                fi.push_back(idx - 1);
            }
            fs.push_back((int)fi.size());
        }
    }
    return finishMesh(v, fs, fi, out);
}
static int readOFF(const std::vector<unsigned char>& buf, OptIO_Mesh* out) {
    std::istringstream in(std::string((const char*)buf.data(), buf.size()));
    std::string magic; in >> magic;
    if (magic != "OFF") return fail(-2, "OFF: missing magic");
    long nv, nf, ne; if (!(in >> nv >> nf >> ne)) return fail(-2, "OFF: bad counts");
    std::vector<float> v((size_t)nv * 3); std::vector<int> fs{ 0 }, fi;
    for (auto& x : v) if (!(in >> x)) return fail(-2, "OFF: truncated vertices");
    for (long f = 0; f < nf; ++f) { int n; if (!(in >> n)) return fail(-2, "OFF: truncated faces"); for (int k = 0; k < n; ++k) { int i; in >> i; fi.push_back(i); } fs.push_back((int)fi.size()); std::string rest; std::getline(in, rest); }
    return finishMesh(v, fs, fi, out);
}
int OptIO_ReadMesh(const char* path, OptIO_Mesh* out) {
    if (!path || !out) return fail(-3, "null argument");
    std::vector<unsigned char> buf;
    if (!readAll(path, buf)) return fail(-1, std::string("cannot read ") + path);
    const std::string e = lowerExt(path);
    if (e == "ply") return readPLY(buf, out);
    if (e == "obj") return readOBJ(buf, out);
    if (e == "off") return readOFF(buf, out);
    return fail(-2, "mesh: unknown extension (expected .ply, .obj or .off)");
}
int OptIO_WritePLY(const char* path, const OptIO_Mesh* m) {
    if (!path || !m) return fail(-3, "null argument");
    FILE* f = fopen(path, "w");
    if (!f) return fail(-1, std::string("cannot write ") + path);
    fprintf(f, "ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n", m->nVertices, m->nFaces);
    for (int i = 0; i < m->nVertices; ++i) fprintf(f, "%.9g %.9g %.9g\n", m->vertices[3 * i], m->vertices[3 * i + 1], m->vertices[3 * i + 2]);
    for (int k = 0; k < m->nFaces; ++k) { fprintf(f, "%d", m->faceStart[k + 1] - m->faceStart[k]); for (int j = m->faceStart[k]; j < m->faceStart[k + 1]; ++j) fprintf(f, " %d", m->faceIndices[j]); fprintf(f, "\n"); }
    fclose(f);
    return 0;
}
void OptIO_FreeMesh(OptIO_Mesh* m) { if (!m) return; free(m->vertices); free(m->faceStart); free(m->faceIndices); m->vertices = nullptr; m->faceStart = m->faceIndices = nullptr; }

int OptIO_MeshNeighbourEdges(const OptIO_Mesh* m, int** head, int** tail, int* nEdges) {
    if (!m || !head || !tail || !nEdges) return fail(-3, "null argument");
    std::vector<std::pair<int, int>> e;
    for (int k = 0; k < m->nFaces; ++k) {
        const int a = m->faceStart[k], n = m->faceStart[k + 1] - a;
        for (int j = 0; j < n; ++j) { const int u = m->faceIndices[a + j], w = m->faceIndices[a + (j + 1) % n]; if (u != w) { e.emplace_back(u, w); e.emplace_back(w, u); } }
    }
    std::sort(e.begin(), e.end()); e.erase(std::unique(e.begin(), e.end()), e.end());   // grouped by head vertex, as the neighbour-list walk produces
    *nEdges = (int)e.size();
    *head = (int*)malloc(sizeof(int) * std::max<size_t>(1, e.size())); *tail = (int*)malloc(sizeof(int) * std::max<size_t>(1, e.size()));
    for (size_t i = 0; i < e.size(); ++i) { (*head)[i] = e[i].first; (*tail)[i] = e[i].second; }
    return 0;
}
void OptIO_FreeInts(int* p) { free(p); }

// ------------------------------------------------------------------------------------------------ text side files
int OptIO_ReadMarkers(const char* path, int maxMarkers, float* positions, float* radii, int* vertexIndices) {
    if (!path) return fail(-3, "null argument");
    std::ifstream in(path);
    if (!in.good()) return fail(-1, std::string("cannot read ") + path);
    unsigned n = 0; if (!(in >> n)) return fail(-2, "markers: bad count");
    for (unsigned m = 0; m < n; ++m) {
        float p[3], r; unsigned vi;
        if (!(in >> p[0] >> p[1] >> p[2] >> r >> vi)) return fail(-2, "markers: truncated");
        if ((int)m < maxMarkers) { if (positions) memcpy(positions + 3 * m, p, sizeof p); if (radii) radii[m] = r; if (vertexIndices) vertexIndices[m] = (int)vi; }
    }
    return (int)n;
}
int OptIO_ReadPairConstraints(const char* path, int maxConstraints, int* rows) {
    if (!path) return fail(-3, "null argument");
    std::ifstream in(path);
    if (!in.good()) return fail(-1, std::string("cannot read ") + path);
    unsigned n = 0; if (!(in >> n)) return fail(-2, "constraints: bad count");
    for (unsigned m = 0; m < n; ++m) {
        int v[4];
        if (!(in >> v[0] >> v[1] >> v[2] >> v[3])) return fail(-2, "constraints: truncated");
        if ((int)m < maxConstraints && rows) memcpy(rows + 4 * m, v, sizeof v);
    }
    return (int)n;
}

}  // extern "C"
