// mesh_io.cpp -- inspection-mesh ingest (replaces osgDB::readNodeFile at
// plan_evaluator/Utility_Functions/src/SceneKeeper.cpp:73-86).
// Triangles come out in file order as 9 floats each; the index of a triangle is its colour ID
// (TriangleData.cpp:23-39).  Formats: binary and ASCII STL, Wavefront OBJ (the OSG plugin's default
// Y-up -> Z-up rotation (x, y, z) -> (x, -z, y) and fan splitting of polygons), and a flat binary
// format (.ippm: "IPPM", uint32 count, 9 floats per triangle) used for synthetic meshes.
#include "mesh_io.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace ipp {

namespace {

std::vector<char> slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::invalid_argument("The following file does not specify a valid 3D model: " + path);
    std::vector<char> buf;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    if (n > 0 && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) {
        fclose(f);
        throw std::invalid_argument("short read on " + path);
    }
    fclose(f);
    return buf;
}

bool has_suffix(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
    return true;
}

void parse_stl(const std::vector<char>& buf, const std::string& path, std::vector<float>& out) {
    if (buf.size() >= 84) {
        uint32_t n;
        memcpy(&n, buf.data() + 80, 4);
        if (buf.size() == 84 + (size_t)n * 50) {  // binary
            out.resize((size_t)n * 9);
            for (uint32_t i = 0; i < n; ++i) memcpy(out.data() + (size_t)i * 9, buf.data() + 84 + (size_t)i * 50 + 12, 36);
            return;
        }
    }
    // ASCII: every "vertex x y z"
    std::string text(buf.begin(), buf.end());
    const char* p = text.c_str();
    while ((p = strstr(p, "vertex")) != nullptr) {
        p += 6;
        for (int k = 0; k < 3; ++k) {
            char* next = nullptr;
            float v = strtof(p, &next);
            if (next == p) throw std::invalid_argument("not a valid STL file: " + path);
            p = next;
            out.push_back(v);
        }
    }
    if (out.empty() || out.size() % 9) throw std::invalid_argument("not a valid STL file: " + path);
}

void parse_obj(const std::vector<char>& buf, const std::string& path, std::vector<float>& out) {
    std::vector<float> verts;  // rotated
    const char* p = buf.data();
    const char* end = p + buf.size();
    std::vector<long> face;
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        std::string line(p, eol);
        p = eol + 1;
        if (line.size() < 2 || !(line[1] == ' ' || line[1] == '\t')) continue;
        if (line[0] == 'v') {
            float x = 0, y = 0, z = 0;
            if (sscanf(line.c_str() + 2, "%f %f %f", &x, &y, &z) != 3) throw std::invalid_argument("bad OBJ vertex in " + path);
            verts.push_back(x);
            verts.push_back(-z);
            verts.push_back(y);
        } else if (line[0] == 'f') {
            face.clear();
            const char* q = line.c_str() + 2;
            while (*q) {
                while (*q && isspace((unsigned char)*q)) ++q;
                if (!*q) break;
                char* next = nullptr;
                long i = strtol(q, &next, 10);
                if (next == q) break;
                long nv = (long)(verts.size() / 3);
                i = i < 0 ? nv + i : i - 1;
                if (i < 0 || i >= nv) throw std::invalid_argument("bad OBJ face index in " + path);
                face.push_back(i);
                q = next;
                while (*q && !isspace((unsigned char)*q)) ++q;  // skip /vt/vn
            }
            for (size_t k = 1; k + 1 < face.size(); ++k) {
                const long tri[3] = {face[0], face[k], face[k + 1]};
                for (long t : tri)
                    for (int c = 0; c < 3; ++c) out.push_back(verts[(size_t)t * 3 + c]);
            }
        }
    }
    if (out.empty()) throw std::invalid_argument("no faces in OBJ file: " + path);
}

}  // namespace

void load_mesh(const std::string& path, std::vector<float>& tris) {
    tris.clear();
    std::vector<char> buf = slurp(path);
    if (has_suffix(path, ".stl")) {
        parse_stl(buf, path, tris);
    } else if (has_suffix(path, ".obj")) {
        parse_obj(buf, path, tris);
    } else if (has_suffix(path, ".ippm")) {
        uint32_t n = 0;
        if (buf.size() < 8 || memcmp(buf.data(), "IPPM", 4)) throw std::invalid_argument("not an IPPM mesh: " + path);
        memcpy(&n, buf.data() + 4, 4);
        if (buf.size() != 8 + (size_t)n * 36) throw std::invalid_argument("truncated IPPM mesh: " + path);
        tris.resize((size_t)n * 9);
        memcpy(tris.data(), buf.data() + 8, (size_t)n * 36);
    } else {
        throw std::invalid_argument("The following file does not specify a valid 3D model: " + path);
    }
    if (tris.size() / 9 > (1u << 24))  // TriangleData.cpp:36
        throw std::logic_error("The input 3D structure has too many primitives (more than 256^3)");
}

}  // namespace ipp
