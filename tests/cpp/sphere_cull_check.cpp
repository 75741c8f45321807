// TEST INFRASTRUCTURE (CPU): the bounding-sphere pre-test of sampleLightUnlessDark (step 0, lighting.cuh /
// lights.cu, transcribed) is conservative: whenever it rejects a light triangle, every sample point of that triangle lies
// below the shading horizon with at least the 1e-3 cosine margin of the per-sample test - checked in double precision on
// random triangles, shading points, normals and view sides, with many configurations close to the horizon.
#include <cmath>
#include <cstdio>
#include <random>
struct V { float x, y, z; };
static V sub(V a, V b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static float dotf(V a, V b) { return std::fmaf(a.z, b.z, std::fmaf(a.y, b.y, a.x * b.x)); }
static float sq(V a) { return dotf(a, a); }
int main() {
    std::mt19937_64 rng(99);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    long fired = 0, cases = 0, violations = 0;
    for (int it = 0; it < 3000000; ++it) {
        const double scale = std::pow(10.0, 1.5 * U(rng));                 // triangle sizes over three decades
        V p = { (float)(5 * U(rng)), (float)(5 * U(rng)), (float)(5 * U(rng)) };
        double nx = U(rng), ny = U(rng), nz = U(rng), nl = std::sqrt(nx * nx + ny * ny + nz * nz);
        if (nl < 1e-3) continue;
        V n = { (float)(nx / nl), (float)(ny / nl), (float)(nz / nl) };   // decoded 16-bit normals are unit to ~1e-4
        n.x *= 1.0f + 1e-4f * (float)U(rng);
        const float vz = (it % 3 == 0 ? -1.0f : 1.0f) * (float)(0.02 + 0.98 * std::fabs(U(rng)));
        // triangle placed around a point whose height above the horizon plane is small in half of the cases
        const double h = (it % 2 ? 0.05 : 3.0) * U(rng), along = 4 * U(rng);
        V base = { (float)(p.x + h * n.x + along * n.y), (float)(p.y + h * n.y - along * n.x), (float)(p.z + h * n.z + 2 * U(rng)) };
        V tri[3];
        for (auto &v : tri) v = { (float)(base.x + scale * U(rng)), (float)(base.y + scale * U(rng)), (float)(base.z + scale * U(rng)) };
        // lights.cu: sphere
        V c = { (tri[0].x + tri[1].x + tri[2].x) * (1.0f / 3.0f), (tri[0].y + tri[1].y + tri[2].y) * (1.0f / 3.0f), (tri[0].z + tri[1].z + tri[2].z) * (1.0f / 3.0f) };
        const float r2 = std::fmax(std::fmax(sq(sub(tri[0], c)), sq(sub(tri[1], c))), sq(sub(tri[2], c)));
        const float r = std::sqrt(r2) * 1.0001f + 1e-30f;
        // lighting.cuh: pre-test
        const V dc = sub(c, p);
        const float t = -(dotf(dc, n) * vz);
        const float margin = t - 1.001f * r * std::fabs(vz);
        const bool cull = r >= 0.0f && margin > 0.0f && margin * margin > 8e-6f * (sq(dc) + r * r) * (vz * vz);
        ++cases;
        if (!cull) continue;
        ++fired;
        for (int s = 0; s < 24; ++s) { // vertices, edge points and interior points
            double a = std::fabs(U(rng)), b = std::fabs(U(rng));
            if (s < 3) { a = s == 0; b = s == 1; }
            if (a + b > 1) { a = 1 - a; b = 1 - b; }
            const double w = 1 - a - b;
            const double qx = a * tri[0].x + b * tri[1].x + w * tri[2].x - p.x, qy = a * tri[0].y + b * tri[1].y + w * tri[2].y - p.y,
                         qz = a * tri[0].z + b * tri[1].z + w * tri[2].z - p.z;
            const double sgn = vz > 0 ? 1.0 : -1.0;
            const double sCos = -(qx * n.x + qy * n.y + qz * n.z) * sgn;
            if (!(sCos > 1e-3 * std::sqrt(qx * qx + qy * qy + qz * qz)))
                ++violations;
        }
    }
    printf("%ld cases, %ld culled, %ld violations\n", cases, fired, violations);
    return violations != 0 || fired < cases / 20;
}
