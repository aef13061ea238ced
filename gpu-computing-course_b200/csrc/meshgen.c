/*
 * meshgen.c — deterministic synthetic triangle meshes for the self-collision path.
 *
 * Workload utilities (NOT part of the collision library and NOT part of the oracle):
 * the same float arrays produced here are handed to the CUDA path, to the CPU
 * oracle and to the reference driver, so all three see bit-identical input.
 *
 * Workloads follow BASELINE.json `configs` / SURVEY.md §8(d):
 *   C1/C2  flag stand-in  (bundled flag-2000-changed.obj is missing from the checkout)
 *   C3     cloth_fold     accordion-folded sheet, dense contacts, shared vertices
 *   C4     soup           random triangle soup, controlled overlap density
 *   C5     two_sheets     two subdivided sheets intersecting along curves
 *
 * Randomness is a counter-based splitmix64 hash keyed by (seed, a, b, c): no state,
 * so any slice of a mesh can be generated independently and in parallel.
 *
 * OBJ writer emits the only dialect the reference parser accepts
 * (/root/reference/CollisionDetection/load_obj.h:50,68): "v %f %f %f" and
 * "f %d/%d %d/%d %d/%d", vertices before faces, trailing newline.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define MG_API __attribute__((visibility("default")))

static inline uint64_t mg_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* uniform double in [0,1) keyed by four counters */
static inline double mg_u01(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    uint64_t h = mg_mix(seed);
    h = mg_mix(h ^ (a * 0xD6E8FEB86659FD93ull));
    h = mg_mix(h ^ (b * 0xA0761D6478BD642Full));
    h = mg_mix(h ^ (c * 0xE7037ED1A0B428DBull));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------ soup (C4) */
/* n triangles, V = 3n private vertices. Centroid uniform in the box
 * [origin, origin+extent); each vertex = centroid + uniform(-h,h)^3. */
MG_API void mg_soup(uint32_t n, double h, uint64_t seed, const double origin[3],
                    const double extent[3], float* xyz, uint32_t* idx) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)n; ++t) {
        double c[3];
        for (int a = 0; a < 3; ++a)
            c[a] = origin[a] + extent[a] * mg_u01(seed, (uint64_t)t, 3, (uint64_t)a);
        for (int v = 0; v < 3; ++v) {
            for (int a = 0; a < 3; ++a) {
                double r = mg_u01(seed, (uint64_t)t, (uint64_t)v, (uint64_t)a);
                xyz[(3 * t + v) * 3 + a] = (float)(c[a] + h * (2.0 * r - 1.0));
            }
            idx[3 * t + v] = (uint32_t)(3 * t + v);
        }
    }
}

/* Grid helper: (nx+1)*(ny+1) vertices already written by the caller; emits
 * 2*nx*ny triangles with vertex offset vbase. Diagonal alternates per quad so
 * the triangulation is not direction-biased. */
static void mg_grid_faces(uint32_t nx, uint32_t ny, uint32_t vbase, uint32_t* idx) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)ny; ++j) {
        for (uint32_t i = 0; i < nx; ++i) {
            uint32_t v00 = vbase + (uint32_t)j * (nx + 1) + i;
            uint32_t v10 = v00 + 1;
            uint32_t v01 = v00 + (nx + 1);
            uint32_t v11 = v01 + 1;
            uint32_t* f = idx + 6 * ((uint64_t)j * nx + i);
            if (((i + j) & 1u) == 0) {
                f[0] = v00; f[1] = v10; f[2] = v11;
                f[3] = v00; f[4] = v11; f[5] = v01;
            } else {
                f[0] = v00; f[1] = v10; f[2] = v01;
                f[3] = v10; f[4] = v11; f[5] = v01;
            }
        }
    }
}

MG_API uint64_t mg_grid_num_verts(uint32_t nx, uint32_t ny) { return (uint64_t)(nx + 1) * (ny + 1); }
MG_API uint64_t mg_grid_num_tris(uint32_t nx, uint32_t ny) { return 2ull * nx * ny; }

/* ------------------------------------------------------------ cloth_fold (C3) */
/* One nx*ny-quad sheet, accordion-folded into `layers` layers along x, layers
 * stacked along y. The layer gap is gap_edges * (mean edge length) and each layer
 * carries a sinusoidal y-perturbation of amplitude amp * gap whose wavelength is
 * wl_edges edges, with a per-layer hashed phase, so adjacent layers interpenetrate
 * densely at every resolution (gap 0.5 edge, amplitude 1.5 gap).
 * Placed inside the reference's hard-coded Morton box (morton.h:43-58):
 * x in [0.2,2.9], y from -0.40 upward, z in [-0.3,1.9]. */
MG_API void mg_cloth_fold(uint32_t nx, uint32_t ny, uint32_t layers, double gap_edges, double amp,
                          double wl_edges, uint64_t seed, float* xyz, uint32_t* idx) {
    const double PI = 3.14159265358979323846;
    const double X0 = 0.2, W = 2.7, Y0 = -0.40, Z0 = -0.3, D = 2.2;
    const double edge = W * (double)layers / (double)nx;
    const double g = gap_edges * edge;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j <= (int64_t)ny; ++j) {
        double v = (double)j / (double)ny;
        for (uint32_t i = 0; i <= nx; ++i) {
            double u = (double)i / (double)nx;
            double s = u * (double)layers;          /* position along the folded strip */
            double l = floor(s);
            if (l >= (double)layers) l = (double)layers - 1.0;
            double f = s - l;                        /* 0..1 inside the layer */
            int odd = ((int)l) & 1;
            double tx = odd ? 1.0 - f : f;           /* triangle wave */
            double ph = 2.0 * PI * mg_u01(seed, (uint64_t)l, 17, 0);
            double ph2 = 2.0 * PI * mg_u01(seed, (uint64_t)l, 17, 1);
            double wob = sin(2.0 * PI * (tx * W / edge) / wl_edges + ph) *
                         cos(2.0 * PI * (double)j / wl_edges + ph2);
            double fade = sin(PI * f);               /* 0 at creases keeps the sheet continuous */
            double x = X0 + W * tx;
            double y = Y0 + g * (s + 0.5) + amp * g * wob * fade;
            double z = Z0 + D * v;
            float* p = xyz + 3 * ((uint64_t)j * (nx + 1) + i);
            p[0] = (float)x; p[1] = (float)y; p[2] = (float)z;
        }
    }
    mg_grid_faces(nx, ny, 0, idx);
}

/* ------------------------------------------------------------ two_sheets (C5) */
/* Two nq*nq-quad sheets over [0.02,0.98]^2 with low-frequency sinusoidal
 * z-displacement around z=0.5 so that they intersect along curves. Unit cube.
 * A small hashed in-plane jitter (<= 0.2 of a cell) de-regularises the grid. */
MG_API void mg_two_sheets(uint32_t nq, uint64_t seed, float* xyz, uint32_t* idx) {
    const double PI = 3.14159265358979323846;
    const double lo = 0.02, span = 0.96;
    const uint64_t vper = (uint64_t)(nq + 1) * (nq + 1);
    const double cell = span / (double)nq;
    for (int s = 0; s < 2; ++s) {
        double p0 = 2.0 * PI * mg_u01(seed, (uint64_t)s, 1, 0);
        double p1 = 2.0 * PI * mg_u01(seed, (uint64_t)s, 1, 1);
        double fx = s ? 3.0 : 2.0, fy = s ? 2.0 : 3.0;
        double a = 0.11;
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j <= (int64_t)nq; ++j) {
            for (uint32_t i = 0; i <= nq; ++i) {
                double jx = (mg_u01(seed, (uint64_t)s + 2, (uint64_t)j, i) - 0.5) * 0.4 * cell;
                double jy = (mg_u01(seed, (uint64_t)s + 4, (uint64_t)j, i) - 0.5) * 0.4 * cell;
                double x = lo + cell * (double)i + ((i > 0 && i < nq) ? jx : 0.0);
                double y = lo + cell * (double)j + ((j > 0 && j < (int64_t)nq) ? jy : 0.0);
                double z = 0.5 + (s ? -0.02 : 0.02) +
                           a * sin(2.0 * PI * fx * x + p0) * cos(2.0 * PI * fy * y + p1);
                float* p = xyz + 3 * (s * vper + (uint64_t)j * (nq + 1) + i);
                p[0] = (float)x; p[1] = (float)y; p[2] = (float)z;
            }
        }
        mg_grid_faces(nq, nq, (uint32_t)(s * vper), idx + 3 * (uint64_t)s * 2ull * nq * nq);
    }
}

/* ----------------------------------------------------------- flag stand-in (C1/C2) */
/* A waving flag inside the reference Morton box. The flag spans x (length) and
 * z (height) and waves in y. Its x-y profile is a trochoid
 *     x(u) = X0 + W*u - a(u,z)*sin(k*u),  y(u) = b*cos(k*u) (+ slow wave)
 * which self-intersects where a*k/W > 1. a(u,z) exceeds that threshold only in
 * `nfold` small (u,z) windows, so the flag interpenetrates itself in a handful of
 * places — the same character as the reference's 20-pair golden result
 * (resources/MyResult.txt). */
MG_API void mg_flag(uint32_t nx, uint32_t nz, uint32_t nfold, uint64_t seed, float* xyz,
                    uint32_t* idx) {
    const double PI = 3.14159265358979323846;
    const double X0 = 0.15, W = 2.75, Z0 = -0.30, D = 2.15;
    const double waves = 9.0;
    const double k = 2.0 * PI * waves;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j <= (int64_t)nz; ++j) {
        double v = (double)j / (double)nz;
        for (uint32_t i = 0; i <= nx; ++i) {
            double u = (double)i / (double)nx;
            double a = 0.55 / k; /* below the self-intersection threshold 1/k (in u units) */
            for (uint32_t f = 0; f < nfold; ++f) {
                double zc = 0.1 + 0.8 * mg_u01(seed, f, 5, 0);
                double wd = 3.0 / (double)nz + 0.002 * mg_u01(seed, f, 5, 1);
                double uc = (1.0 + floor((waves - 1.0) * mg_u01(seed, f, 5, 2))) / waves; /* a crest: cos(k*uc) = 1 */
                double tv = (v - zc) / wd, tu = (u - uc) * waves / 0.6;
                a += (0.80 / k) * exp(-tv * tv - tu * tu);
            }
            double x = X0 + W * (u - a * sin(k * u));
            double y = -0.10 + 0.10 * cos(k * u) * (0.3 + 0.7 * u) +
                       0.08 * sin(2.0 * PI * (1.5 * u + 0.7 * v));
            double z = Z0 + D * v + 0.02 * sin(2.0 * PI * 2.0 * u);
            float* p = xyz + 3 * ((uint64_t)j * (nx + 1) + i);
            p[0] = (float)x; p[1] = (float)y; p[2] = (float)z;
        }
    }
    mg_grid_faces(nx, nz, 0, idx);
}

/* ------------------------------------------------------------------ OBJ writer */
/* "%.9g" round-trips every float through the reference's sscanf("%f"). */
MG_API int mg_write_obj(const char* path, const float* xyz, uint32_t nverts,
                        const uint32_t* idx, uint32_t ntris) {
    FILE* fp = fopen(path, "w");
    if (!fp) return -1;
    static char buf[1 << 20];
    setvbuf(fp, buf, _IOFBF, sizeof buf);
    fprintf(fp, "# b200cd synthetic mesh: %u vertices, %u triangles\n", nverts, ntris);
    for (uint32_t i = 0; i < nverts; ++i)
        fprintf(fp, "v %.9g %.9g %.9g\n", xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    for (uint32_t t = 0; t < ntris; ++t)
        fprintf(fp, "f %u/%u %u/%u %u/%u\n", idx[3 * t] + 1, idx[3 * t] + 1, idx[3 * t + 1] + 1,
                idx[3 * t + 1] + 1, idx[3 * t + 2] + 1, idx[3 * t + 2] + 1);
    int rc = ferror(fp) ? -2 : 0;
    if (fclose(fp) != 0) rc = -3;
    return rc;
}
