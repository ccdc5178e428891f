// ipp_oracle.cpp -- CPU restatement of the reference plan-fitness evaluator.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
// inspection_path_planning_b200/ and libipp.so) may include, link or call this
// file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker and as the host-core baseline.
//
// PARITY UNPINNED: the reference (kaiolae/Inspection-Path-Planning) ships no
// tests, golden vectors or fixtures, and it cannot be built here (it needs
// OpenSceneGraph, Boost, SWIG, Python 2 and an OpenGL context).
// The visibility step of the reference is an OpenGL item-buffer render whose
// arithmetic lives in a GL driver that is not in /root/reference.  This file
// restates the in-tree arithmetic operation by operation (fp64, no FMA
// contraction: build with -ffp-contract=off) and replaces the GL render by a
// software item buffer with the published OpenGL semantics: pixel-centre
// sampling, depth test LESS with primitives submitted in ID order, clip to
// planar depth [zNear, zFar], both faces drawn.
//
// Citations are file:line below /root/reference/ with
//   PE/ = plan_evaluator/Evaluator/src/   UF/ = plan_evaluator/Utility_Functions/src/
//
// Build: see oracle/Makefile (g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace {

// ----------------------------------------------------------------------------
// Constants: UF/Constants.h:19-60
// ----------------------------------------------------------------------------
const double NUM_VIEWPOINTS_INSIDE_INSPECTION_VOLUME = 1000.0;
const double BOX_SAFETY_MARGIN = 2.0;
const double BOX_PADDING = 4.0;
const double MIN_DIST_FROM_BOTTOM = 3.0;
const bool DOWNWARD_CAMERA_ACTIVE = true;
const double translationEnergyFactor = 0.1;
const double rotationEnergyFactor = 1;
const double maxEnergyMultiplier = 1.0;
const double COLLISION_PENALTY_MODIFIER = 2.0;
const double EDGE_SAFETY_BUFFER = 1.5;
const double CAMERA_FAR_PLANE_DIST = 10;

// Depth resolution of the item buffer.  The reference renders into a fixed-point GL depth buffer with depth
// function LESS and submits the triangles in ID order (UF/TriangleData.cpp:93-106), so fragments whose depths
// fall into the same depth-buffer step are ties and the first one drawn -- the lowest ID -- keeps the pixel.
// Restated here as: among the fragments of a pixel whose planar depth is within a relative 2^-16 of the nearest
// one, the lowest ID wins.  (Exactly coplanar, differently triangulated surfaces -- common in CAD exports such
// as mockup_only.stl -- then resolve deterministically instead of by the rounding noise of the last bit.)
const double DEPTH_TIE_REL = 1.0 / 65536.0;

// ----------------------------------------------------------------------------
// osg::Vec3d / osg::Vec3f arithmetic, restated (upstream OSG semantics, SURVEY
// Appendix B): component-wise ops, dot = x*x' + y*y' + z*z' left to right,
// normalize() multiplies by 1/length.
// ----------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
struct V3f {
    float x, y, z;
};
inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 widen(const V3f& a) { return V3{(double)a.x, (double)a.y, (double)a.z}; }
inline V3f narrow(const V3& a) { return V3f{(float)a.x, (float)a.y, (float)a.z}; }
inline V3 operator+(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(const V3& a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(const V3& a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(const V3& a, double s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double length(const V3& a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline void normalize(V3& a) {
    double n = length(a);
    if (n > 0.0) {
        double inv = 1.0 / n;
        a.x *= inv;
        a.y *= inv;
        a.z *= inv;
    }
}
inline bool equal(const V3& a, const V3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// osg::Plane (double): set(p1,p2,p3), flip, distance.  Appendix B.4.
struct Plane {
    double a, b, c, d;
};
inline Plane plane_from_points(const V3& v1, const V3& v2, const V3& v3_) {
    V3 norm = cross(v2 - v1, v3_ - v2);
    double len = length(norm);
    if (len > 1e-6)
        norm = norm / len;
    else
        norm = v3(0, 0, 0);
    return Plane{norm.x, norm.y, norm.z, -(dot(v1, norm))};
}
// UF/OsgHelpers.cpp:284-292 getPlane
inline Plane get_plane(const V3& p1, const V3& p2, const V3& p3, const V3& dir) {
    Plane p = plane_from_points(p1, p2, p3);
    double dn = p.a * dir.x + p.b * dir.y + p.c * dir.z;
    if (dn < 0) {
        p.a = -p.a;
        p.b = -p.b;
        p.c = -p.c;
        p.d = -p.d;
    }
    return p;
}
inline double plane_dist(const Plane& p, const V3& v) { return p.a * v.x + p.b * v.y + p.c * v.z + p.d; }

// Triangle-vs-convex-polytope predicate standing in for
// osgUtil::PolytopeIntersector::containsIntersections() (UF/OsgHelpers.cpp:249-252,
// :323-326; OSG itself is not in /root/reference).  Exact statement: the
// triangle intersects {x : dist_k(x) >= 0 for all k}.  Realised as
// Sutherland-Hodgman clipping in fp64 with two shortcuts; the CUDA kernel runs
// the identical operation sequence so the flags agree bit for bit.
bool polytope_hits_triangle(const Plane* pl, int np, const V3& A, const V3& B, const V3& C) {
    bool all_in = true;
    for (int k = 0; k < np; ++k) {
        double da = plane_dist(pl[k], A), db = plane_dist(pl[k], B), dc = plane_dist(pl[k], C);
        if (da < 0 && db < 0 && dc < 0) return false;
        if (!(da >= 0 && db >= 0 && dc >= 0)) all_in = false;
    }
    if (all_in) return true;
    V3 poly[16];
    V3 out[16];
    double dist[16];
    int n = 3;
    poly[0] = A;
    poly[1] = B;
    poly[2] = C;
    for (int k = 0; k < np; ++k) {
        for (int i = 0; i < n; ++i) dist[i] = plane_dist(pl[k], poly[i]);
        int m = 0;
        for (int i = 0; i < n; ++i) {
            int j = (i + 1 == n) ? 0 : i + 1;
            bool in_i = dist[i] >= 0, in_j = dist[j] >= 0;
            if (in_i) out[m++] = poly[i];
            if (in_i != in_j) {
                double t = dist[i] / (dist[i] - dist[j]);
                out[m++] = poly[i] + (poly[j] - poly[i]) * t;
            }
        }
        n = m;
        if (n == 0) return false;
        for (int i = 0; i < n; ++i) poly[i] = out[i];
    }
    return true;
}

// UF/OsgHelpers.cpp:195-244 generatePolytopeFromPlanEdge (plane order as p.add)
void edge_prism(const V3& edgeStart, const V3& edgeEnd, Plane out[6]) {
    V3 edgeDirection = edgeEnd - edgeStart;
    normalize(edgeDirection);
    V3 up = v3(0, 0, 1);
    if (edgeEnd.x == edgeStart.x && edgeEnd.y == edgeStart.y) up = v3(1, 0, 0);
    V3 right = cross(edgeDirection, up);
    const double B = EDGE_SAFETY_BUFFER;
    // corners are osg::Vec3 (float) in the reference: rounded, then widened
    V3 sLL = widen(narrow(edgeStart - (up * B) - (right * B)));
    V3 sLR = widen(narrow(edgeStart - (up * B) + (right * B)));
    V3 sUL = widen(narrow(edgeStart + (up * B) - (right * B)));
    V3 sUR = widen(narrow(edgeStart + (up * B) + (right * B)));
    V3 eLL = widen(narrow(edgeEnd - (up * B) - (right * B)));
    V3 eLR = widen(narrow(edgeEnd - (up * B) + (right * B)));
    V3 eUL = widen(narrow(edgeEnd + (up * B) - (right * B)));
    V3 eUR = widen(narrow(edgeEnd + (up * B) + (right * B)));
    out[0] = get_plane(eLL, eLR, eUR, -edgeDirection);  // far
    out[1] = get_plane(sLR, sLL, sUL, edgeDirection);   // near
    out[2] = get_plane(sLL, sUL, eUL, right);           // left
    out[3] = get_plane(sLR, sUR, eUR, -right);          // right
    out[4] = get_plane(sUL, sUR, eUR, -up);             // top
    out[5] = get_plane(sLR, sLL, eLL, up);              // bottom
}

// UF/OsgHelpers.cpp:295-318 generateBoxPolytope
void box_polytope(const V3& c, double s, Plane out[6]) {
    V3 r = v3(0, 1, 0), f = v3(1, 0, 0), u = v3(0, 0, 1);
    out[0] = get_plane(c + (f * s), c + (f * s) + r, c + (f * s) + u, -(f * s));      // far
    out[1] = get_plane(c - (f * s), c - (f * s) + r, c - (f * s) + u, (f * s));       // near
    out[2] = get_plane(c - (r * s), c + f - (r * s), c + u - (r * s), (r * s));       // left
    out[3] = get_plane(c + (r * s), c + f + (r * s), c + u + (r * s), -(r * s));      // right
    out[4] = get_plane(c + (u * s), c + f + (u * s), c + (u * s) - r, -(u * s));      // top
    out[5] = get_plane(c - (u * s), c + f - (u * s), c - (u * s) - r, (u * s));       // bottom
}

// ----------------------------------------------------------------------------
// Mesh ingest.  osgDB::readNodeFile is third party:
// binary/ASCII STL in file order; OBJ with the plugin's default Y-up -> Z-up
// rotation (x,y,z)->(x,-z,y) and fan split of polygons.
// ----------------------------------------------------------------------------
struct Mesh {
    std::vector<V3f> v;  // 3 per triangle, file order
};

bool ends_with_ci(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (tolower(s[s.size() - n + i]) != suf[i]) return false;
    return true;
}

void load_stl(const std::string& path, Mesh& m) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::invalid_argument("cannot open " + path);
    std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (buf.size() >= 84) {
        uint32_t n;
        memcpy(&n, &buf[80], 4);
        if (buf.size() == 84 + (size_t)n * 50) {
            m.v.resize((size_t)n * 3);
            for (uint32_t i = 0; i < n; ++i) {
                const char* p = &buf[84 + (size_t)i * 50 + 12];
                float c[9];
                memcpy(c, p, 36);
                for (int k = 0; k < 3; ++k) m.v[(size_t)i * 3 + k] = V3f{c[3 * k], c[3 * k + 1], c[3 * k + 2]};
            }
            return;
        }
    }
    // ASCII STL
    std::istringstream ss(std::string(buf.begin(), buf.end()));
    std::string tok;
    while (ss >> tok) {
        if (tok == "vertex") {
            float x, y, z;
            ss >> x >> y >> z;
            m.v.push_back(V3f{x, y, z});
        }
    }
    if (m.v.empty() || m.v.size() % 3) throw std::invalid_argument("not a valid STL: " + path);
}

void load_obj(const std::string& path, Mesh& m) {
    std::ifstream f(path);
    if (!f) throw std::invalid_argument("cannot open " + path);
    std::vector<V3f> verts;
    std::string line;
    while (std::getline(f, line)) {
        if (line.size() < 2) continue;
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
            std::istringstream ss(line.substr(2));
            float x, y, z;
            ss >> x >> y >> z;
            verts.push_back(V3f{x, -z, y});  // Y-up -> Z-up
        } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
            std::istringstream ss(line.substr(2));
            std::string tok;
            std::vector<long> idx;
            while (ss >> tok) {
                long i = strtol(tok.c_str(), nullptr, 10);
                if (i < 0) i = (long)verts.size() + i; else i = i - 1;
                if (i < 0 || i >= (long)verts.size()) throw std::invalid_argument("bad OBJ index in " + path);
                idx.push_back(i);
            }
            for (size_t k = 1; k + 1 < idx.size(); ++k) {
                m.v.push_back(verts[idx[0]]);
                m.v.push_back(verts[idx[k]]);
                m.v.push_back(verts[idx[k + 1]]);
            }
        }
    }
    if (m.v.empty()) throw std::invalid_argument("no faces in OBJ: " + path);
}

// own flat format used by tests/bench for synthetic meshes: "IPPM" u32 T, 9 floats per triangle
void load_ippm(const std::string& path, Mesh& m) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::invalid_argument("cannot open " + path);
    char magic[4];
    uint32_t n;
    f.read(magic, 4);
    f.read((char*)&n, 4);
    if (!f || memcmp(magic, "IPPM", 4)) throw std::invalid_argument("not an IPPM mesh: " + path);
    m.v.resize((size_t)n * 3);
    f.read((char*)m.v.data(), (std::streamsize)n * 36);
    if (!f) throw std::invalid_argument("truncated IPPM mesh: " + path);
}

void load_mesh(const std::string& path, Mesh& m) {
    if (ends_with_ci(path, ".stl")) load_stl(path, m);
    else if (ends_with_ci(path, ".obj")) load_obj(path, m);
    else if (ends_with_ci(path, ".ippm")) load_ippm(path, m);
    else throw std::invalid_argument("The following file does not specify a valid 3D model: " + path);
}

// ----------------------------------------------------------------------------
// Moore-neighbour tracing, restated from PE/ContourTracing.cpp:70-252
// (input indexed [x][y]; `width` = outer size; result pairs (pos/(w+2)-1, pos%(w+2)-1)).
// ----------------------------------------------------------------------------
std::vector<std::pair<int, int>> moore_trace(const std::vector<std::vector<bool>>& img) {
    int width = (int)img.size();
    int height = (int)img[0].size();
    int W2 = width + 2;
    std::vector<char> padded((size_t)(height + 2) * W2, 0), border((size_t)(height + 2) * W2, 0);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) padded[(x + 1) + (y + 1) * W2] = img[x][y];
    std::vector<std::pair<int, int>> visited;
    bool inside = false;
    const int nb[8][2] = {{-1, 7}, {-3 - width, 7}, {-width - 2, 1}, {-1 - width, 1},
                          {1, 3},  {3 + width, 3},  {width + 2, 5},  {1 + width, 5}};
    for (int y = 0; y < height + 2; ++y) {
        for (int x = 0; x < W2; ++x) {
            int pos = x + y * W2;
            if (border[pos] && !inside) {
                inside = false;
            } else if (padded[pos] && inside) {
                continue;
            } else if (!padded[pos] && inside) {
                inside = true;
            } else if (padded[pos] && !inside) {
                visited.push_back({pos / W2 - 1, pos % W2 - 1});
                border[pos] = 1;
                int checkLocationNr = 1, startPos = pos, counter = 0, counter2 = 0;
                while (true) {
                    int checkPosition = pos + nb[checkLocationNr - 1][0];
                    int newCheckLocationNr = nb[checkLocationNr - 1][1];
                    if (padded[checkPosition]) {
                        visited.push_back({checkPosition / W2 - 1, checkPosition % W2 - 1});
                        if (checkPosition == startPos) {
                            counter++;
                            if (newCheckLocationNr == 1 || counter >= 3) {
                                inside = true;
                                break;
                            }
                        }
                        checkLocationNr = newCheckLocationNr;
                        pos = checkPosition;
                        counter2 = 0;
                        border[checkPosition] = 1;
                    } else {
                        checkLocationNr = 1 + (checkLocationNr % 8);
                        if (counter2 > 8) {
                            counter2 = 0;
                            break;
                        } else {
                            counter2++;
                        }
                    }
                }
                // NB: like the reference, `pos` was advanced by the trace; the
                // outer loop recomputes it from (x, y) on the next iteration.
            }
        }
    }
    return visited;
}

// std::ostream << double with default precision (6 significant digits): UF/HelperMethods.h:194-218
std::string vec_to_string(const double* v, int n, int numElements) {
    std::ostringstream oss;
    if (n > 0 && numElements <= n) {
        if (numElements == n) {
            for (int i = 0; i + 1 < n; ++i) oss << v[i] << ",";
            oss << v[n - 1];
        } else if (numElements == 1) {
            oss << v[0];
        } else {
            for (int i = 0; i < numElements; ++i) oss << v[i] << ",";
        }
    }
    return oss.str();
}

typedef std::vector<uint32_t> Bits;
inline void bits_set(Bits& b, size_t i) { b[i >> 5] |= (1u << (i & 31)); }
inline bool bits_get(const Bits& b, size_t i) { return (b[i >> 5] >> (i & 31)) & 1u; }

struct Oracle {
    // scene (UF/SceneKeeper.cpp:65-71, UF/TriangleData.cpp:41-72)
    Mesh mesh;
    size_t T = 0;
    std::vector<double> area;
    double totalArea = 0;
    V3f bbmin, bbmax;
    double collisionPenalty = 0;
    V3 sceneCenter;
    // canonical (lexicographically sorted) vertex order per triangle, used only
    // by the item buffer so that exact duplicate faces produce exact depth ties
    std::vector<V3f> cv;
    // triangles whose fp64 normal (B-A)x(C-A) is exactly zero produce no fragments (GL: zero-area primitive)
    std::vector<char> degenerate;
    // sensor (UF/CameraEstimator.cpp:53-63)
    double sampleInterval, fovY, aspect, zNear, zFar;
    unsigned imgH, imgW;
    bool frontCam, downCam;
    // interpreter
    bool postProcessing = false, hasStart = false, loops = false;
    V3 start;
    std::vector<V3> boxCenters;
    std::vector<std::vector<std::vector<int>>> boxVolume;  // [z][x][y]
    std::vector<std::vector<std::vector<double>>> singleLevel, complete;
    double maxAllowedEnergy = -1;
    std::map<std::string, Bits> memo;
    // counters for the CPU baseline
    uint64_t nSnapshots = 0, nPixels = 0, nCollisionQueries = 0;
    // scratch item buffer
    std::vector<double> zbuf;
    std::vector<int32_t> idbuf;
    std::vector<uint32_t> glDepth;  // depth model 1: 24-bit window depth per pixel
    std::vector<double> z1;  // edge-epsilon classification: per pixel and sample position
    std::vector<int32_t> idWide, idNarrow;
    size_t W32() const { return (T + 31) / 32; }

    // ---- setup ----
    void init(const std::string& path, const double* specs, bool post, const double* startXYZ, bool loopsAround) {
        load_mesh(path, mesh);
        T = mesh.v.size() / 3;
        if (T > (1u << 24)) throw std::logic_error("more than 256^3 primitives");  // UF/TriangleData.cpp:36
        area.resize(T);
        cv.resize(T * 3);
        degenerate.assign(T, 0);
        bbmin = V3f{INFINITY, INFINITY, INFINITY};
        bbmax = V3f{-INFINITY, -INFINITY, -INFINITY};
        for (size_t i = 0; i < T; ++i) {
            V3 v1 = widen(mesh.v[3 * i]), v2 = widen(mesh.v[3 * i + 1]), v3_ = widen(mesh.v[3 * i + 2]);
            double s1 = length(v2 - v1), s2 = length(v3_ - v1), s3 = length(v3_ - v2);
            double hp = (s1 + s2 + s3) / 2.0;
            area[i] = std::sqrt(hp * (hp - s1) * (hp - s2) * (hp - s3));  // Heron, UF/TriangleData.cpp:56-61
            V3f c[3] = {mesh.v[3 * i], mesh.v[3 * i + 1], mesh.v[3 * i + 2]};
            std::sort(c, c + 3, [](const V3f& a, const V3f& b) {
                if (a.x != b.x) return a.x < b.x;
                if (a.y != b.y) return a.y < b.y;
                return a.z < b.z;
            });
            {
                V3 n = cross(widen(c[1]) - widen(c[0]), widen(c[2]) - widen(c[0]));
                degenerate[i] = (n.x == 0 && n.y == 0 && n.z == 0);
            }
            for (int k = 0; k < 3; ++k) {
                cv[3 * i + k] = c[k];
                const V3f& p = mesh.v[3 * i + k];
                bbmin.x = std::min(bbmin.x, p.x); bbmin.y = std::min(bbmin.y, p.y); bbmin.z = std::min(bbmin.z, p.z);
                bbmax.x = std::max(bbmax.x, p.x); bbmax.y = std::max(bbmax.y, p.y); bbmax.z = std::max(bbmax.z, p.z);
            }
        }
        totalArea = 0;
        for (size_t i = 0; i < T; ++i) totalArea += area[i];  // UF/TriangleData.cpp:65-72
        // osg::BoundingBox is float: extents are float subtractions (UF/OsgHelpers.cpp:499-511)
        double xl = std::fabs(bbmax.x - bbmin.x), yl = std::fabs(bbmax.y - bbmin.y), zl = std::fabs(bbmax.z - bbmin.z);
        collisionPenalty = COLLISION_PENALTY_MODIFIER * std::max(zl, std::max(xl, yl));  // UF/SceneKeeper.cpp:67
        // bounding-sphere centre of a single-geode scene = float bbox centre (Appendix B.3)
        V3f c = V3f{(bbmin.x + bbmax.x) * 0.5f, (bbmin.y + bbmax.y) * 0.5f, (bbmin.z + bbmax.z) * 0.5f};
        sceneCenter = widen(c);

        sampleInterval = specs[0];
        double fov_x = specs[1], fov_y = specs[2];
        // argument swap: ImageViewerCaptureTool(fovY:=fov_x, fovX:=fov_y, ...) UF/CameraEstimator.cpp:62,
        // UF/ImageViewerCaptureTool.cpp:22-26
        fovY = fov_x;
        aspect = fov_y / fov_x;
        imgH = (unsigned)specs[3];
        imgW = (unsigned)(imgH * aspect);
        zNear = specs[4];
        zFar = specs[5];
        frontCam = specs[6] != 0;
        downCam = specs[7] != 0;
        zbuf.resize((size_t)imgW * imgH);
        idbuf.resize((size_t)imgW * imgH);

        postProcessing = post;
        loops = loopsAround;
        if (startXYZ) {
            hasStart = true;
            start = v3(startXYZ[0], startXYZ[1], startXYZ[2]);
        }
        divide_into_boxes();
        if (!postProcessing) {
            produce_simple_circling_plans();
            // PE/PlanEnergyEvaluator.cpp:40-46
            std::vector<std::vector<std::vector<double>>> sp = complete_circling_plans();
            if (sp.empty()) throw std::logic_error("no circling plans could be generated");
            maxAllowedEnergy = decode_energy(sp.front()) * maxEnergyMultiplier;
        } else {
            maxAllowedEnergy = -1;
        }
    }

    bool polytope_hits_mesh(const Plane* pl) {
        nCollisionQueries++;
        for (size_t i = 0; i < T; ++i)
            if (polytope_hits_triangle(pl, 6, widen(mesh.v[3 * i]), widen(mesh.v[3 * i + 1]), widen(mesh.v[3 * i + 2])))
                return true;
        return false;
    }

    // PE/PlanInterpreterBoxOrder.cpp:193-251
    void divide_into_boxes() {
        double padded = ((bbmax.x - bbmin.x) + 2 * BOX_PADDING) * ((bbmax.y - bbmin.y) + 2 * BOX_PADDING) *
                        ((bbmax.z - bbmin.z) + BOX_PADDING);
        double viewpointVolume = padded / NUM_VIEWPOINTS_INSIDE_INSPECTION_VOLUME;
        double interval = std::cbrt(viewpointVolume);
        double minZ = bbmin.z + MIN_DIST_FROM_BOTTOM;
        double maxZ = bbmax.z;
        if (DOWNWARD_CAMERA_ACTIVE) maxZ += BOX_PADDING;
        for (double z = minZ; z < maxZ; z += interval) {
            std::vector<std::vector<int>> layer;
            for (double x = bbmin.x - BOX_PADDING; x < bbmax.x + BOX_PADDING; x += interval) {
                std::vector<int> row;
                for (double y = bbmin.y - BOX_PADDING; y < bbmax.y + BOX_PADDING; y += interval) {
                    V3 vp = v3(x, y, z);
                    Plane pl[6];
                    box_polytope(vp, BOX_SAFETY_MARGIN, pl);
                    if (polytope_hits_mesh(pl)) {
                        row.push_back(-1);
                        continue;
                    }
                    box_polytope(vp, CAMERA_FAR_PLANE_DIST, pl);
                    if (polytope_hits_mesh(pl)) {
                        row.push_back((int)boxCenters.size());
                        boxCenters.push_back(vp);
                    } else {
                        row.push_back(-2);
                    }
                }
                layer.push_back(row);
            }
            boxVolume.push_back(layer);
        }
    }

    // PE/PlanInterpreterBoxOrder.cpp:24-38.  NB the reference indexes the layer
    // as currentlayer[y][x] although it was filled [x][y]; kept verbatim.
    bool box_neighbors_object(int x, int y, int z) const {
        const std::vector<std::vector<int>>& cur = boxVolume[z];
        for (int xn = -1; xn < 2; ++xn)
            for (int yn = -1; yn < 2; ++yn) {
                if (x + xn < 0 || (unsigned)(x + xn) >= cur[0].size() || y + yn < 0 || (unsigned)(y + yn) >= cur.size())
                    continue;
                if (cur[y + yn][x + xn] == -1) return true;
            }
        return false;
    }

    // PE/PlanInterpreterBoxOrder.cpp:40-109
    void produce_simple_circling_plans() {
        for (unsigned z = 0; z < boxVolume.size(); ++z) {
            const std::vector<std::vector<int>>& cur = boxVolume[z];
            std::vector<std::vector<bool>> edgeBoxes;
            for (unsigned y = 0; y < cur.size(); ++y) {
                std::vector<bool> row;
                for (unsigned x = 0; x < cur[y].size(); ++x) row.push_back(cur[y][x] == -1 || box_neighbors_object(x, y, z));
                edgeBoxes.push_back(row);
            }
            std::vector<std::pair<int, int>> traces = moore_trace(edgeBoxes);
            if (traces.empty()) continue;
            std::vector<std::vector<double>> plan;
            for (size_t i = 0; i < traces.size(); ++i) {
                int x = traces[i].first, y = traces[i].second;
                auto has = [&](int a, int b) { return std::find(traces.begin(), traces.end(), std::make_pair(a, b)) != traces.end(); };
                bool nextX = has(x + 1, y), prevX = has(x - 1, y), nextY = has(x, y + 1), prevY = has(x, y - 1);
                if (!((prevX && nextX) || (prevY && nextY))) plan.push_back(std::vector<double>{(double)boxVolume[z][y][x]});
            }
            singleLevel.push_back(plan);
        }
    }

    // PE/PlanInterpreterBoxOrder.cpp:111-160
    std::vector<std::vector<std::vector<double>>> complete_circling_plans() {
        if (!complete.empty()) return complete;
        int total = (int)singleLevel.size();
        for (int step = 1; step <= total; ++step) {
            std::vector<std::vector<double>> cur;
            std::vector<int> levels;
            for (int lv = 0; lv < total; lv += step) levels.push_back(lv);
            int skew;
            int mid = levels[levels.size() / 2];
            double middle = ((double)total) / 2.0;
            if (levels.size() % 2 == 1) {
                skew = int(middle - mid);
            } else {
                int second = levels[(levels.size() / 2) - 1];
                double med = double(mid + second) / 2.0;
                skew = int(middle - med);
            }
            for (int& l : levels) l += skew;
            for (int l : levels) {
                // the reference indexes without a range check; out-of-range is UB there, an error here
                if (l < 0 || l >= total) throw std::logic_error("circling plan level out of range");
                cur.insert(cur.end(), singleLevel[l].begin(), singleLevel[l].end());
            }
            complete.push_back(cur);
        }
        return complete;
    }

    // ---- decode (PE/PlanInterpreterBoxOrder.cpp:253-324) ----
    V3 box_position(size_t id) const {
        if (id >= boxCenters.size()) throw std::out_of_range("box id out of range");  // assert :328
        return boxCenters[id];
    }
    static size_t gene_id(double g) { return (size_t)(g + 0.5); }

    // UF/OsgHelpers.cpp:421-446
    V3 viewing_direction(const V3& a, const V3& b) const {
        V3 d;
        if (a.x == b.x && a.y == b.y) {
            d = sceneCenter - ((b + a) / 2.0);
            d.z = 0;
        } else {
            d = cross(v3(0, 0, 1), b - a);
            V3 mid = b + (b - a) / 2.0;
            V3 want = sceneCenter - mid;
            want.z = 0;
            if (dot(d, want) < 0) d = -d;
        }
        return d;
    }
    // UF/OsgHelpers.cpp:451-489 (calculateViewingDirectionTowardsPrimitives): the same axis as viewing_direction, but the side is
    // chosen by looking -- both candidate headings are rendered along the edge and the one with the lower coverage score wins; a
    // tie falls back to viewing_direction.  The candidates are passed on unnormalised, as the reference does.
    V3 viewing_direction_towards_primitives(const V3& a, const V3& b) {
        V3 d;
        if (a.x == b.x && a.y == b.y) {
            d = sceneCenter - ((b + a) / 2.0);
            d.z = 0;
        } else {
            d = cross(v3(0, 0, 1), b - a);
        }
        V3 r = -d;
        Bits c1(W32(), 0), c2(W32(), 0);
        colors_during_traversal(a, b, d, c1);
        double s1 = coverage_score(c1);
        colors_during_traversal(a, b, r, c2);
        double s2 = coverage_score(c2);
        if (s2 == s1) return viewing_direction(a, b);
        return s2 > s1 ? d : r;
    }
    // PE/PlanInterpreterBoxOrder.cpp:258-276
    V3 sensor_direction(const V3& prev, const V3& next, float offset) {
        V3 d = postProcessing ? viewing_direction_towards_primitives(prev, next) : viewing_direction(prev, next);
        normalize(d);
        double rad = std::atan2(d.y, d.x);
        rad += offset;
        d = v3(std::cos(rad), std::sin(rad), 0);
        normalize(d);
        return d;
    }
    void interpret_one(const V3* prev, const double* gene, int width, std::vector<V3>& pos, std::vector<V3>& dirs) {
        size_t id = gene_id(gene[0]);
        float off = 0;
        if (width > 1) off = (float)gene[1];
        V3 cur = box_position(id);
        pos.push_back(cur);
        if (prev) dirs.push_back(sensor_direction(*prev, cur, off));
    }
    void interpret_plan(const std::vector<std::vector<double>>& plan, std::vector<V3>& pos, std::vector<V3>& dirs) {
        if (plan.empty()) return;
        if (hasStart) pos.push_back(start);
        for (const auto& g : plan) {
            if (!pos.empty()) {
                V3 prev = pos.back();
                interpret_one(&prev, g.data(), (int)g.size(), pos, dirs);
            } else {
                interpret_one(nullptr, g.data(), (int)g.size(), pos, dirs);
            }
        }
    }

    // ---- energy (UF/OsgHelpers.cpp:105-163, :246-262) ----
    bool edge_collides(const V3& a, const V3& b) {
        Plane pl[6];
        edge_prism(a, b, pl);
        return polytope_hits_mesh(pl);
    }
    double energy_one_move(const V3& s, const V3& e, const V3* prevPoint, bool* collided) {
        if (collided) *collided = false;
        if (equal(s, e)) return 0.0;
        double energy = translationEnergyFactor * length(e - s);
        if (prevPoint) {
            V3 in = s - *prevPoint, out = e - s;
            double angleCos = dot(in, out) / (length(in) * length(out));
            double turn = 1.0 - angleCos;
            energy += turn * rotationEnergyFactor;
        }
        if (collisionPenalty <= 0) return energy;
        if (edge_collides(s, e)) {
            energy += collisionPenalty;
            if (collided) *collided = true;
        }
        return energy;
    }
    double energy_of_positions(const std::vector<V3>& p, double* pathLen = nullptr, int* nColl = nullptr) {
        if (pathLen) *pathLen = 0;
        if (nColl) *nColl = 0;
        if (p.empty()) return 0;
        double e = 0;
        for (size_t k = 1; k < p.size(); ++k) {
            bool c = false;
            e += energy_one_move(p[k - 1], p[k], k > 1 ? &p[k - 2] : nullptr, &c);
            if (pathLen) *pathLen += length(p[k] - p[k - 1]);
            if (nColl && c) (*nColl)++;
        }
        return e;
    }
    double decode_energy(const std::vector<std::vector<double>>& plan, double* pathLen = nullptr, int* nColl = nullptr) {
        std::vector<V3> pos, dirs;
        interpret_plan(plan, pos, dirs);
        return energy_of_positions(pos, pathLen, nColl);
    }

    // ---- sensor (UF/CameraEstimator.cpp:83-97, :131-178) ----
    void sampling_positions(const V3& a, const V3& b, std::vector<V3f>& out, double interval) const {
        if (interval <= 0) interval = sampleInterval;
        V3 mv = b - a;
        out.push_back(narrow(a));
        V3 next = a + (mv * interval) / double(length(mv));
        while (length(next - a) < length(mv)) {
            out.push_back(narrow(next));
            next = next + (mv * interval) / length(mv);
        }
        out.push_back(narrow(b));
    }

    // Software item buffer for one camera pose; replaces viewer.frame() +
    // glReadPixels + getAllColorsInFrame (UF/ImageViewerCaptureTool.cpp:59-70,
    // UF/CameraEstimator.cpp:99-129).  `seen` receives the nominal seen set.
    // With strictSeen/lenientSeen non-null the pose is additionally evaluated at
    // the pixel centre and at four sample positions displaced by +-eps_px
    // pixels, with a relative depth tolerance eps_depth, to build the
    // edge-epsilon classification used by the parity tests:
    //   strict  : the triangle wins some pixel at all five positions, clearly in front
    //   lenient : the triangle wins (or nearly ties) some pixel at any of the five
    // so that strict <= nominal <= lenient, and a correct evaluator working in
    // lower precision must return a set between strict and lenient.
    struct Margins {
        double eps_px = 0.01;     // sample displacement in pixels
        double eps_depth = 1e-5;  // relative depth margin
    } margins;
    // Depth model of the item buffer.  0 (default): planar depth, fragments within a relative 2^-16 of the nearest are ties, lowest
    // ID wins (DEPTH_TIE_REL above).  1: the OpenGL depth buffer as the reference's viewer most likely had it (parity is unpinned:
    // the GL driver is not in the reference tree) -- window depth z_w = f / (f - n) * (1 - n / z) of the perspective projection
    // with n = zNear, f = zFar (UF/ImageViewerCaptureTool.cpp:22-27, no automatic near/far, :35-36), quantised to an unsigned
    // 24-bit value, depth function GL_LESS with the triangles submitted in ID order (UF/TriangleData.cpp:93-106), cleared to 1.0,
    // and the top-left fill rule for pixel centres exactly on an edge.  Used to bound what the choice of model can change
    // (tests/test_oracle_known_answers.py, tools/depth_model_report.py); classification sets are only built for model 0.
    int depthModel = 0;

    struct Cam {
        V3 eye, f, s, u;
        double tanX, tanY;
        int W, H;
    };
    struct Setup {
        double na[3], nb[3], nc[3], vol;
        int x0, x1, y0, y1;
        bool ok;
    };
    Cam make_cam(const V3& eye, const V3& center, const V3& up) const {
        Cam c;
        c.eye = eye;
        // osg::Matrixd::makeLookAt (Appendix B.2)
        c.f = center - eye;
        normalize(c.f);
        c.s = cross(c.f, up);
        normalize(c.s);
        c.u = cross(c.s, c.f);
        normalize(c.u);
        const double PI = 3.14159265358979323846;
        c.tanY = std::tan((fovY * 0.5) * PI / 180.0);  // osg makePerspective
        c.tanX = c.tanY * aspect;
        c.W = (int)imgW;
        c.H = (int)imgH;
        return c;
    }
    void setup_tri(const Cam& cam, size_t i, Setup& S) const {
        S.ok = false;
        if (degenerate[i]) return;
        const V3 &s = cam.s, &u = cam.u, &f = cam.f;
        const int W = cam.W, H = cam.H;
        V3 A = widen(cv[3 * i]) - cam.eye, B = widen(cv[3 * i + 1]) - cam.eye, C = widen(cv[3 * i + 2]) - cam.eye;
        // camera coordinates (x along s, y along u, z = planar depth along f)
        V3 a = v3(dot(A, s), dot(A, u), dot(A, f)), b = v3(dot(B, s), dot(B, u), dot(B, f)), c = v3(dot(C, s), dot(C, u), dot(C, f));
        double lo = zNear * (1 - 1e-3), hi = zFar * (1 + 1e-3);
        if (a.z < lo && b.z < lo && c.z < lo) return;
        if (a.z > hi && b.z > hi && c.z > hi) return;
        // conservative pixel rectangle: clip to z >= lo, project
        V3 poly[4];
        int n = 0;
        const V3 t[3] = {a, b, c};
        for (int k = 0; k < 3; ++k) {
            const V3& p = t[k];
            const V3& q = t[(k + 1) % 3];
            bool ip = p.z >= lo, iq = q.z >= lo;
            if (ip) poly[n++] = p;
            if (ip != iq) {
                double tt = (lo - p.z) / (q.z - p.z);
                poly[n++] = p + (q - p) * tt;
                poly[n - 1].z = lo;
            }
        }
        if (n == 0) return;
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        for (int k = 0; k < n; ++k) {
            double X = poly[k].x / poly[k].z, Y = poly[k].y / poly[k].z;
            xmin = std::min(xmin, X); xmax = std::max(xmax, X);
            ymin = std::min(ymin, Y); ymax = std::max(ymax, Y);
        }
        // pixel i centre: X = ((i+0.5)/W*2-1)*tanX
        double fx0 = (xmin / cam.tanX + 1.0) * 0.5 * W - 0.5, fx1 = (xmax / cam.tanX + 1.0) * 0.5 * W - 0.5;
        double fy0 = (ymin / cam.tanY + 1.0) * 0.5 * H - 0.5, fy1 = (ymax / cam.tanY + 1.0) * 0.5 * H - 0.5;
        if (fx1 < -2 || fy1 < -2 || fx0 > W + 1 || fy0 > H + 1) return;
        S.x0 = std::max(0, (int)std::floor(fx0) - 1);
        S.x1 = std::min(W - 1, (int)std::ceil(fx1) + 1);
        S.y0 = std::max(0, (int)std::floor(fy0) - 1);
        S.y1 = std::min(H - 1, (int)std::ceil(fy1) + 1);
        if (S.x0 > S.x1 || S.y0 > S.y1) return;
        // homogeneous edge functions e_a = D.(b x (c-b)) etc. with D = (X, Y, 1); depth t = vol / (e_a+e_b+e_c)
        V3 Na = cross(b, c - b), Nb = cross(c, a - c), Nc = cross(a, b - a);
        V3 Ng = cross(b - a, c - a);
        S.na[0] = Na.x; S.na[1] = Na.y; S.na[2] = Na.z;
        S.nb[0] = Nb.x; S.nb[1] = Nb.y; S.nb[2] = Nb.z;
        S.nc[0] = Nc.x; S.nc[1] = Nc.y; S.nc[2] = Nc.z;
        S.vol = dot(a, Ng);
        S.ok = true;
    }
    // coverage + depth of one triangle at image-plane position (X, Y); returns false if not covered
    static inline bool frag(const Setup& S, double X, double Y, double& t) {
        double eu = S.na[0] * X + S.na[1] * Y + S.na[2];
        double ev = S.nb[0] * X + S.nb[1] * Y + S.nb[2];
        double ew = S.nc[0] * X + S.nc[1] * Y + S.nc[2];
        double det = eu + ev + ew;
        bool in = ((eu >= 0 && ev >= 0 && ew >= 0) || (eu <= 0 && ev <= 0 && ew <= 0)) && det != 0;
        if (!in) return false;
        t = S.vol / det;
        return true;
    }

    // the same with the triangle dilated by r (image-plane units): every edge function, normalised by its gradient, may be short
    // of zero by r.  Used for the lenient class only: a sliver thinner than the sample displacement can lie between the five
    // sample positions and still be the rightful owner of a pixel whose centre is within rounding of its edge.
    static inline bool frag_dilated(const Setup& S, double X, double Y, double r, double& t) {
        double eu = S.na[0] * X + S.na[1] * Y + S.na[2];
        double ev = S.nb[0] * X + S.nb[1] * Y + S.nb[2];
        double ew = S.nc[0] * X + S.nc[1] * Y + S.nc[2];
        double gu = std::sqrt(S.na[0] * S.na[0] + S.na[1] * S.na[1]), gv = std::sqrt(S.nb[0] * S.nb[0] + S.nb[1] * S.nb[1]),
               gw = std::sqrt(S.nc[0] * S.nc[0] + S.nc[1] * S.nc[1]);
        double det = eu + ev + ew;
        if (det == 0) return false;
        bool pos = eu >= -r * gu && ev >= -r * gv && ew >= -r * gw, neg = eu <= r * gu && ev <= r * gv && ew <= r * gw;
        if (!pos && !neg) return false;
        t = S.vol / det;
        return true;
    }

    // coverage with the top-left fill rule: a pixel centre exactly on an edge belongs to the triangle only if the edge is a left edge
    // or a top edge (image y up), so that of two triangles that share the edge exactly one owns the pixel.  Edge functions are
    // oriented so that the inside is positive.
    static inline bool frag_top_left(const Setup& S, double X, double Y, double& t) {
        double e[3] = {S.na[0] * X + S.na[1] * Y + S.na[2], S.nb[0] * X + S.nb[1] * Y + S.nb[2], S.nc[0] * X + S.nc[1] * Y + S.nc[2]};
        const double* N[3] = {S.na, S.nb, S.nc};
        double det = e[0] + e[1] + e[2];
        if (det == 0) return false;
        const double sg = det > 0 ? 1.0 : -1.0;  // all three edge functions carry the sign of det inside the triangle
        for (int k = 0; k < 3; ++k) {
            const double v = sg * e[k];
            if (v < 0) return false;
            if (v == 0) {
                // the inward normal of the edge in the image plane is sg * (N[0], N[1]): a left edge has it pointing towards +x, a top
                // edge (horizontal, interior below it) towards -y
                const double nx = sg * N[k][0], ny = sg * N[k][1];
                if (!(nx > 0 || (nx == 0 && ny < 0))) return false;
            }
        }
        t = S.vol / det;
        return true;
    }

    // Columns [xa, xb] of image row Y (inside [x0, x1]) outside which no pixel centre can pass the coverage test of frag() or
    // frag_dilated(): a pure accelerator of the loops below (long thin triangles have bounding boxes hundreds of times their
    // area), the per-pixel tests are unchanged.  Conservative: every edge function may miss its sign by `slack` image-plane units
    // times its gradient plus a rounding allowance, and the interval is widened by a pixel on either side.
    static inline bool row_span(const Setup& S, const Cam& cam, double Y, double slack, int x0, int x1, int& xa, int& xb) {
        const double* N[3] = {S.na, S.nb, S.nc};
        double loP = -INFINITY, hiP = INFINITY, loN = -INFINITY, hiN = INFINITY;
        bool okP = true, okN = true;
        for (int k = 0; k < 3; ++k) {
            const double n0 = N[k][0], n1y = N[k][1] * Y, c = n1y + N[k][2];
            const double g = std::sqrt(n0 * n0 + N[k][1] * N[k][1]);
            const double m = slack * g + 1e-9 * (std::fabs(n0) * cam.tanX + std::fabs(n1y) + std::fabs(N[k][2]));
            if (n0 > 0) {
                loP = std::max(loP, (-m - c) / n0);
                hiN = std::min(hiN, (m - c) / n0);
            } else if (n0 < 0) {
                hiP = std::min(hiP, (-m - c) / n0);
                loN = std::max(loN, (m - c) / n0);
            } else {
                if (c < -m) okP = false;
                if (c > m) okN = false;
            }
        }
        if (!(loP <= hiP)) okP = false;
        if (!(loN <= hiN)) okN = false;
        if (!okP && !okN) return false;
        const double lo = std::min(okP ? loP : INFINITY, okN ? loN : INFINITY), hi = std::max(okP ? hiP : -INFINITY, okN ? hiN : -INFINITY);
        const double fa = (lo / cam.tanX + 1.0) * 0.5 * cam.W - 0.5, fb = (hi / cam.tanX + 1.0) * 0.5 * cam.W - 0.5;
        xa = !(fa > x0) ? x0 : (fa >= x1 + 2 ? x1 + 1 : std::max(x0, (int)std::floor(fa) - 1));
        xb = !(fb < x1) ? x1 : (fb <= x0 - 2 ? x0 - 1 : std::min(x1, (int)std::ceil(fb) + 1));
        return xa <= xb;
    }

    void render(const V3& eye, const V3& center, const V3& up, Bits& seen, Bits* strictSeen, Bits* lenientSeen) {
        nSnapshots++;
        nPixels += (uint64_t)imgW * imgH;
        Cam cam = make_cam(eye, center, up);
        const int W = cam.W, H = cam.H;
        size_t NP = (size_t)W * H;
        std::fill(zbuf.begin(), zbuf.begin() + NP, INFINITY);
        std::fill(idbuf.begin(), idbuf.begin() + NP, -1);
        Setup S;
        if (depthModel == 1) {
            // OpenGL-style: one pass in ID order, 24-bit window depth, LESS
            const double fn = zFar / (zFar - zNear);
            const double dmax = 16777215.0;  // 2^24 - 1
            std::vector<uint32_t>& dbuf = glDepth;
            dbuf.assign(NP, 16777215u);
            for (size_t i = 0; i < T; ++i) {
                setup_tri(cam, i, S);
                if (!S.ok) continue;
                for (int py = S.y0; py <= S.y1; ++py) {
                    double Y = ((py + 0.5) / H * 2.0 - 1.0) * cam.tanY;
                    int xa, xb;
                    if (!row_span(S, cam, Y, 0.0, S.x0, S.x1, xa, xb)) continue;
                    for (int px = xa; px <= xb; ++px) {
                        double X = ((px + 0.5) / W * 2.0 - 1.0) * cam.tanX;
                        double t;
                        if (!frag_top_left(S, X, Y, t)) continue;
                        if (!(t >= zNear && t <= zFar)) continue;  // clip to the view volume
                        const double zw = fn * (1.0 - zNear / t);
                        const uint32_t d = (uint32_t)std::floor(std::min(std::max(zw, 0.0), 1.0) * dmax + 0.5);
                        size_t pix = (size_t)py * W + px;
                        if (d < dbuf[pix]) {
                            dbuf[pix] = d;
                            idbuf[pix] = (int32_t)i;
                        }
                    }
                }
            }
            for (size_t p = 0; p < NP; ++p)
                if (idbuf[p] >= 0) bits_set(seen, (size_t)idbuf[p]);
            return;
        }
        // pass 1: nearest depth per pixel; pass 2: lowest ID within the depth-tie band of the nearest (see DEPTH_TIE_REL)
        for (int pass = 0; pass < 2; ++pass) {
            for (size_t i = 0; i < T; ++i) {
                setup_tri(cam, i, S);
                if (!S.ok) continue;
                for (int py = S.y0; py <= S.y1; ++py) {
                    double Y = ((py + 0.5) / H * 2.0 - 1.0) * cam.tanY;
                    int xa, xb;
                    if (!row_span(S, cam, Y, 0.0, S.x0, S.x1, xa, xb)) continue;
                    for (int px = xa; px <= xb; ++px) {
                        double X = ((px + 0.5) / W * 2.0 - 1.0) * cam.tanX;
                        double t;
                        if (!frag(S, X, Y, t)) continue;
                        if (!(t >= zNear && t <= zFar)) continue;
                        size_t pix = (size_t)py * W + px;
                        if (pass == 0) {
                            if (t < zbuf[pix]) zbuf[pix] = t;
                        } else if (idbuf[pix] < 0 && t <= zbuf[pix] * (1.0 + DEPTH_TIE_REL)) {
                            idbuf[pix] = (int32_t)i;  // ascending ID: the first one inside the band is the lowest
                        }
                    }
                }
            }
        }
        for (size_t p = 0; p < NP; ++p)
            if (idbuf[p] >= 0) bits_set(seen, (size_t)idbuf[p]);
        if (!strictSeen) return;

        // ---- edge-epsilon classification ----
        const int K = 5;
        const double ex = margins.eps_px * 2.0 * cam.tanX / W, ey = margins.eps_px * 2.0 * cam.tanY / H;
        const double offx[K] = {0, ex, ex, -ex, -ex}, offy[K] = {0, ey, -ey, ey, -ey};
        const double d = margins.eps_depth;
        const double lenLo = zNear * (1 - d), lenHi = zFar * (1 + d), strLo = zNear * (1 + d), strHi = zFar * (1 - d);
        const double bandWide = (1.0 + DEPTH_TIE_REL) * (1 + d), bandNarrow = (1.0 + DEPTH_TIE_REL) * (1 - d);
        // z1: nearest depth per sample position; idWide / idNarrow: lowest ID inside the widened / narrowed tie band
        z1.assign(NP * K, INFINITY);
        idWide.assign(NP * K, -1);
        idNarrow.assign(NP * K, -1);
        // every sample position and the dilated test lie within this distance (image-plane units) of the pixel centre's own test
        const double spanSlack = std::max(ex, ey) * 1.5 + ex + ey;
        for (int pass = 0; pass < 2; ++pass) {
            for (size_t i = 0; i < T; ++i) {
                setup_tri(cam, i, S);
                if (!S.ok) continue;
                for (int py = S.y0; py <= S.y1; ++py) {
                    double Y = ((py + 0.5) / H * 2.0 - 1.0) * cam.tanY;
                    int xa, xb;
                    if (!row_span(S, cam, Y, spanSlack, S.x0, S.x1, xa, xb)) continue;
                    for (int px = xa; px <= xb; ++px) {
                        double X = ((px + 0.5) / W * 2.0 - 1.0) * cam.tanX;
                        size_t pix = ((size_t)py * W + px) * K;
                        if (pass == 1) {
                            // lenient, continuum form: the pixel centre within eps_px of the triangle and the depth inside the
                            // widened band of the nearest surface there
                            double t;
                            if (frag_dilated(S, X, Y, std::max(ex, ey) * 1.5, t) && t >= lenLo && t <= lenHi && t <= z1[pix] * bandWide)
                                bits_set(*lenientSeen, i);
                        }
                        for (int k = 0; k < K; ++k) {
                            double t;
                            if (!frag(S, X + offx[k], Y + offy[k], t)) continue;
                            if (t < lenLo || t > lenHi) continue;
                            if (pass == 0) {
                                if (t < z1[pix + k]) z1[pix + k] = t;
                            } else {
                                if (idWide[pix + k] < 0 && t <= z1[pix + k] * bandWide) idWide[pix + k] = (int32_t)i;
                                if (idNarrow[pix + k] < 0 && t <= z1[pix + k] * bandNarrow && t >= strLo && t <= strHi)
                                    idNarrow[pix + k] = (int32_t)i;
                                // lenient: inside the widened band of any sample position (ID competition ignored)
                                if (t <= z1[pix + k] * bandWide) bits_set(*lenientSeen, i);
                            }
                        }
                    }
                }
            }
        }
        // strict: the same triangle certainly owns the pixel at all five sample positions
        for (size_t p = 0; p < NP; ++p) {
            int32_t id = idWide[p * K];
            if (id < 0) continue;
            bool ok = true;
            for (int k = 0; k < K && ok; ++k) ok = idWide[p * K + k] == id && idNarrow[p * K + k] == id;
            if (ok) bits_set(*strictSeen, (size_t)id);
        }
    }

    // UF/CameraEstimator.cpp:131-178
    void colors_during_traversal(const V3& cur, const V3& next, const V3& heading, Bits& seen, Bits* strictSeen = nullptr,
                                 Bits* lenientSeen = nullptr) {
        std::vector<V3f> samples;
        if (equal(cur, next)) samples.push_back(narrow(cur));
        else sampling_positions(cur, next, samples, sampleInterval);
        V3f hf = narrow(heading);
        for (const V3f& p : samples) {
            if (frontCam) {
                // point+robotHeading is an osg::Vec3 (float) sum, :151
                V3f c = V3f{p.x + hf.x, p.y + hf.y, p.z + hf.z};
                render(widen(p), widen(c), v3(0, 0, 1), seen, strictSeen, lenientSeen);
            }
            if (downCam) {
                V3f c = V3f{p.x - 0.0f, p.y - 0.0f, p.z - 1.0f};  // point-UP_VECTOR in float, :162
                render(widen(p), widen(c), heading, seen, strictSeen, lenientSeen);
            }
        }
    }

    // ---- evaluation (PE/PlanCoverageEstimator.cpp) ----
    // :43-118
    Bits evaluate_memoised(const std::vector<std::vector<double>>& plan) {
        V3 prevLoc = v3(0, 0, 0);
        std::string prevName = "None";
        if (hasStart) {
            prevName = "start";
            prevLoc = start;
        }
        Bits observed(W32(), 0);
        for (size_t i = 0; i < plan.size(); ++i) {
            const std::vector<double>& part = plan[i];
            if (i != 0) prevName = vec_to_string(plan[i - 1].data(), (int)plan[i - 1].size(), 1);
            std::string curName = vec_to_string(part.data(), (int)part.size(), (int)part.size());
            std::string key = prevName + "-" + curName;
            auto it = memo.find(key);
            if (it == memo.end()) {
                std::vector<V3> pos, dirs;
                if (prevName != "start" && prevName != "None") prevLoc = box_position(gene_id(plan[i - 1][0]));
                if (prevName == "None") {
                    interpret_one(nullptr, part.data(), (int)part.size(), pos, dirs);
                } else {
                    pos.push_back(prevLoc);
                    interpret_one(&prevLoc, part.data(), (int)part.size(), pos, dirs);
                }
                if (pos.size() < 2) continue;
                Bits cur(W32(), 0);
                for (size_t e = 0; e < dirs.size(); ++e) {
                    colors_during_traversal(pos[e], pos[e + 1], dirs[e], cur);
                    for (size_t w = 0; w < observed.size(); ++w) observed[w] |= cur[w];
                    memo[key] = cur;
                }
            } else {
                for (size_t w = 0; w < observed.size(); ++w) observed[w] |= it->second[w];
            }
        }
        return observed;
    }
    // :128-151
    Bits evaluate_internal(const std::vector<std::vector<double>>& plan) {
        std::vector<V3> pos, dirs;
        interpret_plan(plan, pos, dirs);
        Bits observed(W32(), 0);
        V3 cur = pos[0];
        for (size_t i = 0; i < dirs.size(); ++i) {
            V3 nxt = pos[i + 1];
            colors_during_traversal(cur, nxt, dirs[i], observed);
            cur = nxt;
        }
        return observed;
    }
    // UF/TriangleData.cpp:155-194
    double coverage_score(const Bits& b) const {
        double cov = 0;
        for (size_t i = 0; i < T; ++i)
            if (bits_get(b, i)) cov += area[i];
        return 1.0 - (cov / totalArea);
    }
    // :201-282; out = {coverage, energy, pathLen, nCollisions}
    void evaluate_plan(const std::vector<std::vector<double>>& plan, bool memoization, bool disableLimit, double out[4]) {
        if (plan.empty()) {
            out[0] = 1.0; out[1] = 0; out[2] = 0; out[3] = 0;
            return;
        }
        std::vector<std::vector<double>> copy = plan;
        if (loops) copy.push_back(plan[0]);
        double pathLen = 0;
        int nColl = 0;
        double energy = decode_energy(copy, &pathLen, &nColl);
        out[1] = energy; out[2] = pathLen; out[3] = nColl;
        bool within = (maxAllowedEnergy == -1) ? true : (energy <= maxAllowedEnergy);  // PE/PlanEnergyEvaluator.cpp:28-36
        if (!disableLimit && !within) {
            out[0] = 1;
            return;
        }
        Bits observed = memoization ? evaluate_memoised(copy) : evaluate_internal(copy);
        out[0] = coverage_score(observed);
    }
    // :284-327
    int update_memoised_edges(const std::vector<std::vector<std::vector<double>>>& pop) {
        std::set<std::string> keep;
        for (const auto& sol : pop) {
            std::string prevName = "start";
            for (const auto& step : sol) {
                std::string curName = vec_to_string(step.data(), (int)step.size(), (int)step.size());
                keep.insert(prevName + "-" + curName);
                prevName = vec_to_string(step.data(), (int)step.size(), 1);
            }
        }
        for (auto it = memo.begin(); it != memo.end();) {
            if (keep.find(it->first) == keep.end()) it = memo.erase(it);
            else ++it;
        }
        return (int)memo.size();
    }
};

thread_local std::string g_err;

std::vector<std::vector<double>> unpack_plan(const double* genes, int n, int width) {
    std::vector<std::vector<double>> p((size_t)n);
    for (int i = 0; i < n; ++i) p[i].assign(genes + (size_t)i * width, genes + (size_t)(i + 1) * width);
    return p;
}

}  // namespace

#define ORACLE_TRY try {
#define ORACLE_CATCH                 \
    }                                \
    catch (const std::exception& e) { \
        g_err = e.what();            \
        return 1;                    \
    }                                \
    return 0;

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

int oracle_create(const char* path, const double* specs, int postProcessing, const double* startXYZ, int loops, void** out) {
    ORACLE_TRY
    Oracle* o = new Oracle();
    try {
        o->init(path, specs, postProcessing != 0, startXYZ, loops != 0);
    } catch (...) {
        delete o;
        throw;
    }
    *out = o;
    ORACLE_CATCH
}
void oracle_destroy(void* h) { delete (Oracle*)h; }
int64_t oracle_num_triangles(void* h) { return (int64_t)((Oracle*)h)->T; }
int oracle_num_boxes(void* h) { return (int)((Oracle*)h)->boxCenters.size(); }
double oracle_max_allowed_energy(void* h) { return ((Oracle*)h)->maxAllowedEnergy; }
double oracle_total_area(void* h) { return ((Oracle*)h)->totalArea; }
double oracle_collision_penalty(void* h) { return ((Oracle*)h)->collisionPenalty; }
void oracle_scene_info(void* h, double* center3, double* bbox6) {
    Oracle* o = (Oracle*)h;
    center3[0] = o->sceneCenter.x; center3[1] = o->sceneCenter.y; center3[2] = o->sceneCenter.z;
    bbox6[0] = o->bbmin.x; bbox6[1] = o->bbmin.y; bbox6[2] = o->bbmin.z;
    bbox6[3] = o->bbmax.x; bbox6[4] = o->bbmax.y; bbox6[5] = o->bbmax.z;
}
void oracle_image_size(void* h, int* w, int* hh) { *w = (int)((Oracle*)h)->imgW; *hh = (int)((Oracle*)h)->imgH; }
void oracle_areas(void* h, double* out) { Oracle* o = (Oracle*)h; memcpy(out, o->area.data(), o->T * sizeof(double)); }
void oracle_vertices(void* h, float* out) { Oracle* o = (Oracle*)h; memcpy(out, o->mesh.v.data(), o->T * 36); }
void oracle_box_centres(void* h, double* out) {
    Oracle* o = (Oracle*)h;
    for (size_t i = 0; i < o->boxCenters.size(); ++i) {
        out[3 * i] = o->boxCenters[i].x; out[3 * i + 1] = o->boxCenters[i].y; out[3 * i + 2] = o->boxCenters[i].z;
    }
}
void oracle_grid_dims(void* h, int* nz, int* nx, int* ny) {
    Oracle* o = (Oracle*)h;
    *nz = (int)o->boxVolume.size();
    *nx = *nz ? (int)o->boxVolume[0].size() : 0;
    *ny = (*nz && *nx) ? (int)o->boxVolume[0][0].size() : 0;
}
void oracle_grid(void* h, int* out) {  // [z][x][y]
    Oracle* o = (Oracle*)h;
    size_t k = 0;
    for (auto& l : o->boxVolume) for (auto& r : l) for (int v : r) out[k++] = v;
}
// simple plans in CSR form: first call with ids == NULL to get sizes
int oracle_simple_plans(void* h, int* nPlans, int* nGenes, double* ids, int* offsets) {
    ORACLE_TRY
    Oracle* o = (Oracle*)h;
    auto sp = o->complete_circling_plans();
    *nPlans = (int)sp.size();
    int tot = 0;
    for (auto& p : sp) tot += (int)p.size();
    *nGenes = tot;
    if (ids) {
        int k = 0;
        offsets[0] = 0;
        for (size_t i = 0; i < sp.size(); ++i) {
            for (auto& g : sp[i]) ids[k++] = g[0];
            offsets[i + 1] = k;
        }
    }
    ORACLE_CATCH
}
int oracle_evaluate_plan(void* h, const double* genes, int n, int width, int memo, int noLimit, double* out4) {
    ORACLE_TRY
    ((Oracle*)h)->evaluate_plan(unpack_plan(genes, n, width), memo != 0, noLimit != 0, out4);
    ORACLE_CATCH
}
// population in CSR form (width-1 or width-2 genes, same width for all); fitness pop x 4
int oracle_evaluate_population(void* h, const double* genes, int width, const int64_t* offsets, int64_t pop, int memo, int noLimit,
                               double* fitness) {
    ORACLE_TRY
    Oracle* o = (Oracle*)h;
    for (int64_t p = 0; p < pop; ++p)
        o->evaluate_plan(unpack_plan(genes + offsets[p] * width, (int)(offsets[p + 1] - offsets[p]), width), memo != 0, noLimit != 0,
                         fitness + 4 * p);
    ORACLE_CATCH
}
int oracle_update_memoised_edges(void* h, const double* genes, int width, const int64_t* offsets, int64_t pop, int* remaining) {
    ORACLE_TRY
    std::vector<std::vector<std::vector<double>>> all;
    for (int64_t p = 0; p < pop; ++p) all.push_back(unpack_plan(genes + offsets[p] * width, (int)(offsets[p + 1] - offsets[p]), width));
    *remaining = ((Oracle*)h)->update_memoised_edges(all);
    ORACLE_CATCH
}
int oracle_memo_size(void* h) { return (int)((Oracle*)h)->memo.size(); }
int oracle_interpret_plan(void* h, const double* genes, int n, int width, double* pos, double* dirs, int* nPos, int* nDirs) {
    ORACLE_TRY
    std::vector<V3> p, d;
    ((Oracle*)h)->interpret_plan(unpack_plan(genes, n, width), p, d);
    *nPos = (int)p.size();
    *nDirs = (int)d.size();
    for (size_t i = 0; i < p.size(); ++i) { pos[3 * i] = p[i].x; pos[3 * i + 1] = p[i].y; pos[3 * i + 2] = p[i].z; }
    for (size_t i = 0; i < d.size(); ++i) { dirs[3 * i] = d[i].x; dirs[3 * i + 1] = d[i].y; dirs[3 * i + 2] = d[i].z; }
    ORACLE_CATCH
}
// seen-triangle bitmap of one edge a->b with heading computed from (a, b, angle offset).
// strict/lenient may be NULL; otherwise they receive the edge-epsilon classification sets.
int oracle_edge_bitmap_xyz(void* h, const double* a, const double* b, double angle, uint32_t* seen, uint32_t* strict_, uint32_t* lenient) {
    ORACLE_TRY
    Oracle* o = (Oracle*)h;
    V3 A = v3(a[0], a[1], a[2]), B = v3(b[0], b[1], b[2]);
    V3 hd = o->sensor_direction(A, B, (float)angle);
    Bits s(o->W32(), 0), st(o->W32(), 0), le(o->W32(), 0);
    o->colors_during_traversal(A, B, hd, s, strict_ ? &st : nullptr, strict_ ? &le : nullptr);
    memcpy(seen, s.data(), s.size() * 4);
    if (strict_) { memcpy(strict_, st.data(), st.size() * 4); memcpy(lenient, le.data(), le.size() * 4); }
    ORACLE_CATCH
}
int oracle_edge_bitmap(void* h, int prev, int curr, double angle, uint32_t* seen, uint32_t* strict_, uint32_t* lenient) {
    ORACLE_TRY
    Oracle* o = (Oracle*)h;
    V3 A = prev < 0 ? o->start : o->box_position((size_t)prev);
    if (prev < 0 && !o->hasStart) throw std::invalid_argument("no start location");
    V3 B = o->box_position((size_t)curr);
    double a[3] = {A.x, A.y, A.z}, b[3] = {B.x, B.y, B.z};
    if (oracle_edge_bitmap_xyz(h, a, b, angle, seen, strict_, lenient)) throw std::runtime_error(g_err);
    ORACLE_CATCH
}
// one camera pose -> per-pixel triangle ids (-1 = background), row-major [y][x]
int oracle_render_ids(void* h, const double* eye, const double* center, const double* up, int32_t* ids) {
    ORACLE_TRY
    Oracle* o = (Oracle*)h;
    Bits s(o->W32(), 0);
    o->render(v3(eye[0], eye[1], eye[2]), v3(center[0], center[1], center[2]), v3(up[0], up[1], up[2]), s, nullptr, nullptr);
    memcpy(ids, o->idbuf.data(), (size_t)o->imgW * o->imgH * 4);
    ORACLE_CATCH
}
int oracle_edge_collides(void* h, const double* a, const double* b) {
    Oracle* o = (Oracle*)h;
    return o->edge_collides(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2])) ? 1 : 0;
}
int oracle_box_hits_mesh(void* h, const double* c, double halfSide) {
    Oracle* o = (Oracle*)h;
    Plane pl[6];
    box_polytope(v3(c[0], c[1], c[2]), halfSide, pl);
    return o->polytope_hits_mesh(pl) ? 1 : 0;
}
void oracle_edge_prism_planes(const double* a, const double* b, double* out24) {
    Plane pl[6];
    edge_prism(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2]), pl);
    for (int k = 0; k < 6; ++k) { out24[4 * k] = pl[k].a; out24[4 * k + 1] = pl[k].b; out24[4 * k + 2] = pl[k].c; out24[4 * k + 3] = pl[k].d; }
}
int oracle_sampling_positions(void* h, const double* a, const double* b, double interval, float* out, int cap) {
    Oracle* o = (Oracle*)h;
    std::vector<V3f> s;
    V3 A = v3(a[0], a[1], a[2]), B = v3(b[0], b[1], b[2]);
    if (equal(A, B)) s.push_back(narrow(A));
    else o->sampling_positions(A, B, s, interval);
    int n = (int)s.size();
    for (int i = 0; i < n && i < cap; ++i) { out[3 * i] = s[i].x; out[3 * i + 1] = s[i].y; out[3 * i + 2] = s[i].z; }
    return n;
}
void oracle_heading(void* h, const double* a, const double* b, double angle, double* out3) {
    Oracle* o = (Oracle*)h;
    V3 d = o->sensor_direction(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2]), (float)angle);
    out3[0] = d.x; out3[1] = d.y; out3[2] = d.z;
}
double oracle_energy_of_positions(void* h, const double* pos, int n, double* pathLen, int* nColl) {
    Oracle* o = (Oracle*)h;
    std::vector<V3> p;
    for (int i = 0; i < n; ++i) p.push_back(v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
    return o->energy_of_positions(p, pathLen, nColl);
}
double oracle_coverage_of_bitmap(void* h, const uint32_t* bits) {
    Oracle* o = (Oracle*)h;
    Bits b(bits, bits + o->W32());
    return o->coverage_score(b);
}
void oracle_set_margins(void* h, double eps_px, double eps_depth) {
    Oracle* o = (Oracle*)h;
    o->margins.eps_px = eps_px;
    o->margins.eps_depth = eps_depth;
}
void oracle_set_depth_model(void* h, int model) { ((Oracle*)h)->depthModel = model; }
void oracle_counters(void* h, uint64_t* out3) {
    Oracle* o = (Oracle*)h;
    out3[0] = o->nSnapshots; out3[1] = o->nPixels; out3[2] = o->nCollisionQueries;
}
int oracle_moore_trace(const unsigned char* img, int nOuter, int nInner, int* outPairs, int cap) {
    std::vector<std::vector<bool>> v((size_t)nOuter, std::vector<bool>((size_t)nInner));
    for (int i = 0; i < nOuter; ++i) for (int j = 0; j < nInner; ++j) v[i][j] = img[i * nInner + j] != 0;
    auto t = moore_trace(v);
    int n = (int)t.size();
    for (int i = 0; i < n && i < cap; ++i) { outPairs[2 * i] = t[i].first; outPairs[2 * i + 1] = t[i].second; }
    return n;
}
}
