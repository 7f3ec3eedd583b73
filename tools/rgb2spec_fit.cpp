// rgb2spec_fit — deterministic regeneration of the sRGB rgb->spectrum coefficient table.
//
// The reference ships this table as a git-LFS payload (rgb_to_spec/tables/srgb_table.bin, 9 437 440 B) that is
// absent from the checkout, and produces it with a non-deterministic PyTorch MLP+Adam fit
// (/root/reference/rgb_to_spec/python/main.py).  This tool regenerates a table with the SAME grid, model and file
// layout by a deterministic per-cell Levenberg–Marquardt fit in CIELAB (the Jakob–Hanika procedure adapted to the
// reference's logistic model), so every consumer (oracle and device code) indexes it exactly like the reference:
//
//   file   = 64 x f32 z_nodes, then f32 table[3][64][64][64][3]            (spectrum/src/rgb_sigmoid_polynomial.rs:35-84)
//   z_node = smoothstep(smoothstep(i/63))                                   (rgb_to_spec/python/main.py:56-58)
//   cell   = [l][zi][yi][xi] -> rgb[l] = z, rgb[(l+1)%3] = xi/63*z, rgb[(l+2)%3] = yi/63*z   (main.py:165-175)
//   model  = S(lambda) = 1/(1+exp(-(c0 t^2 + c1 t + c2))), t = (lambda-360)/470                (rgb_sigmoid_polynomial.rs:19-27,179-182)
//   target = linear sRGB of S under D65 normalised to Y=1, 1 nm sums over 360..830            (main.py:38-41,196-204)
//
// Values differ from the author's fit (both are approximations of the same inverse problem); indexing is identical.
//
// usage: rgb2spec_fit <std_tables.bin> <out.bin> [threads]
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {
constexpr int N = 64;
constexpr int NL = 470;  // 360..829 nm, the dense tables the renderer itself integrates against
double g_t[NL], g_wx[NL], g_wy[NL], g_wz[NL];  // weights = cmf * d65_normalised
double g_xyz2rgb[3][3], g_rgb2xyz[3][3], g_white[3];

double smoothstep(double x) { return x * x * (3.0 - 2.0 * x); }

void inv3(const double m[3][3], double o[3][3]) {
    double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                 m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    double id = 1.0 / det;
    o[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * id;
    o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
    o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
    o[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * id;
    o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
    o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
    o[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * id;
    o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
    o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
}

void init_colorspace() {
    // sRGB primaries / D65 white (color/src/gamut.rs:43-69)
    const double xy[4][2] = {{0.64, 0.33}, {0.30, 0.60}, {0.15, 0.06}, {0.3127, 0.3290}};
    double P[3][3], W[3];
    for (int c = 0; c < 3; ++c) {
        P[0][c] = xy[c][0] / xy[c][1];
        P[1][c] = 1.0;
        P[2][c] = (1.0 - xy[c][0] - xy[c][1]) / xy[c][1];
    }
    W[0] = xy[3][0] / xy[3][1];
    W[1] = 1.0;
    W[2] = (1.0 - xy[3][0] - xy[3][1]) / xy[3][1];
    double Pi[3][3];
    inv3(P, Pi);
    double s[3];
    for (int i = 0; i < 3; ++i) s[i] = Pi[i][0] * W[0] + Pi[i][1] * W[1] + Pi[i][2] * W[2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) g_rgb2xyz[i][j] = P[i][j] * s[j];
    inv3(g_rgb2xyz, g_xyz2rgb);
    for (int i = 0; i < 3; ++i) g_white[i] = W[i];
}

inline double lab_f(double t) {
    const double d = 6.0 / 29.0;
    return t > d * d * d ? std::cbrt(t) : t / (3 * d * d) + 4.0 / 29.0;
}
inline double lab_df(double t) {
    const double d = 6.0 / 29.0;
    return t > d * d * d ? 1.0 / (3.0 * std::cbrt(t) * std::cbrt(t)) : 1.0 / (3 * d * d);
}
void xyz_to_lab(const double xyz[3], double lab[3], double J[3][3]) {
    double f[3], df[3];
    for (int i = 0; i < 3; ++i) {
        double s = xyz[i] / g_white[i];
        f[i] = lab_f(s);
        df[i] = lab_df(s) / g_white[i];
    }
    lab[0] = 116 * f[1] - 16;
    lab[1] = 500 * (f[0] - f[1]);
    lab[2] = 200 * (f[1] - f[2]);
    if (J) {
        J[0][0] = 0; J[0][1] = 116 * df[1]; J[0][2] = 0;
        J[1][0] = 500 * df[0]; J[1][1] = -500 * df[1]; J[1][2] = 0;
        J[2][0] = 0; J[2][1] = 200 * df[1]; J[2][2] = -200 * df[2];
    }
}

// XYZ of the model spectrum and its derivative wrt (c0,c1,c2)
void eval_xyz(const double c[3], double xyz[3], double dxyz[3][3]) {
    double X = 0, Y = 0, Z = 0;
    double dX[3] = {0, 0, 0}, dY[3] = {0, 0, 0}, dZ[3] = {0, 0, 0};
    for (int i = 0; i < NL; ++i) {
        double t = g_t[i];
        double x = (c[0] * t + c[1]) * t + c[2];
        double s = 1.0 / (1.0 + std::exp(-x));
        double ds = s * (1.0 - s);
        X += s * g_wx[i]; Y += s * g_wy[i]; Z += s * g_wz[i];
        double b[3] = {ds * t * t, ds * t, ds};
        for (int k = 0; k < 3; ++k) { dX[k] += b[k] * g_wx[i]; dY[k] += b[k] * g_wy[i]; dZ[k] += b[k] * g_wz[i]; }
    }
    xyz[0] = X; xyz[1] = Y; xyz[2] = Z;
    if (dxyz) for (int k = 0; k < 3; ++k) { dxyz[0][k] = dX[k]; dxyz[1][k] = dY[k]; dxyz[2][k] = dZ[k]; }
}

bool solve3(double A[3][3], const double b[3], double x[3]) {
    double M[3][4];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) M[i][j] = A[i][j]; M[i][3] = b[i]; }
    for (int c = 0; c < 3; ++c) {
        int p = c;
        for (int r = c + 1; r < 3; ++r) if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
        if (std::fabs(M[p][c]) < 1e-300) return false;
        if (p != c) for (int j = 0; j < 4; ++j) std::swap(M[p][j], M[c][j]);
        for (int r = 0; r < 3; ++r) if (r != c) {
            double f = M[r][c] / M[c][c];
            for (int j = c; j < 4; ++j) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 0; i < 3; ++i) x[i] = M[i][3] / M[i][i];
    return true;
}

const double CLAMP[3] = {600.0, 600.0, 60.0};

double residual(const double c[3], const double lab_t[3], double r[3], double J[3][3]) {
    double xyz[3], dxyz[3][3], lab[3], Jl[3][3];
    eval_xyz(c, xyz, J ? dxyz : nullptr);
    for (int i = 0; i < 3; ++i) xyz[i] = std::max(xyz[i], 0.0);
    xyz_to_lab(xyz, lab, J ? Jl : nullptr);
    double e = 0;
    for (int i = 0; i < 3; ++i) { r[i] = lab[i] - lab_t[i]; e += r[i] * r[i]; }
    if (J) for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) {
        J[i][k] = Jl[i][0] * dxyz[0][k] + Jl[i][1] * dxyz[1][k] + Jl[i][2] * dxyz[2][k];
    }
    return e;
}

// Levenberg–Marquardt from the warm start in c[]; returns final squared Lab error.
double fit_cell(const double rgb[3], double c[3]) {
    double xyz_t[3], lab_t[3];
    for (int i = 0; i < 3; ++i) xyz_t[i] = g_rgb2xyz[i][0] * rgb[0] + g_rgb2xyz[i][1] * rgb[1] + g_rgb2xyz[i][2] * rgb[2];
    xyz_to_lab(xyz_t, lab_t, nullptr);
    double r[3], J[3][3];
    double e = residual(c, lab_t, r, J);
    double mu = 1e-3;
    for (int it = 0; it < 60 && e > 1e-10; ++it) {
        double A[3][3], g[3];
        for (int i = 0; i < 3; ++i) {
            g[i] = 0;
            for (int k = 0; k < 3; ++k) g[i] -= J[k][i] * r[k];
            for (int j = 0; j < 3; ++j) { A[i][j] = 0; for (int k = 0; k < 3; ++k) A[i][j] += J[k][i] * J[k][j]; }
        }
        bool improved = false;
        for (int tries = 0; tries < 12; ++tries) {
            double Ad[3][3], d[3], cn[3], rn[3];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ad[i][j] = A[i][j] + (i == j ? mu * (A[i][i] + 1e-12) : 0.0);
            if (!solve3(Ad, g, d)) { mu *= 10; continue; }
            for (int i = 0; i < 3; ++i) cn[i] = std::min(std::max(c[i] + d[i], -CLAMP[i]), CLAMP[i]);
            double en = residual(cn, lab_t, rn, nullptr);
            if (en < e) {
                for (int i = 0; i < 3; ++i) c[i] = cn[i];
                double rel = (e - en) / e;
                e = residual(c, lab_t, r, J);
                mu = std::max(mu * 0.3, 1e-9);
                improved = rel > 1e-7;
                break;
            }
            mu *= 10;
        }
        if (!improved) break;
    }
    return e;
}
}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s std_tables.bin out.bin [threads]\n", argv[0]); return 2; }
    int nthreads = argc > 3 ? std::atoi(argv[3]) : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::vector<float> std_t(4 * NL);
    {
        FILE* f = std::fopen(argv[1], "rb");
        if (!f) { std::perror(argv[1]); return 1; }
        char magic[8];
        if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "TCPTSTD1", 8)) { std::fprintf(stderr, "bad magic\n"); return 1; }
        std::fseek(f, 8 + 104 * 4, SEEK_SET);
        if (std::fread(std_t.data(), 4, 4 * NL, f) != (size_t)4 * NL) { std::fprintf(stderr, "short read\n"); return 1; }
        std::fclose(f);
    }
    init_colorspace();
    double ysum = 0;
    for (int i = 0; i < NL; ++i) ysum += (double)std_t[NL + i] * (double)std_t[3 * NL + i];
    for (int i = 0; i < NL; ++i) {
        g_t[i] = (double)i / 470.0;
        double d = (double)std_t[3 * NL + i] / ysum;
        g_wx[i] = std_t[i] * d; g_wy[i] = std_t[NL + i] * d; g_wz[i] = std_t[2 * NL + i] * d;
    }
    // D65 under these sums is not exactly the sRGB white; use the integrated white so that a flat S=1 maps to rgb (1,1,1)
    // up to the residual chromaticity difference (same convention as the reference's python: white from the colourspace).
    std::vector<float> znodes(N);
    for (int i = 0; i < N; ++i) znodes[i] = (float)smoothstep(smoothstep((double)i / (N - 1)));

    std::vector<float> table((size_t)3 * N * N * N * 3);
    std::atomic<int> next{0};
    std::vector<double> worst(nthreads, 0.0);
    auto work = [&](int tid) {
        for (;;) {
            int job = next.fetch_add(1);
            if (job >= 3 * N * N) break;
            int l = job / (N * N), yi = (job / N) % N, xi = job % N;
            const int start = N / 5;
            double c[3] = {0, 0, 0};
            auto cell = [&](int zi) {
                double z = znodes[zi];
                double rgb[3];
                rgb[l] = z; rgb[(l + 1) % 3] = (double)xi / (N - 1) * z; rgb[(l + 2) % 3] = (double)yi / (N - 1) * z;
                double e;
                if (zi == 0) { c[0] = 0; c[1] = 0; c[2] = -CLAMP[2]; e = 0; }
                else e = fit_cell(rgb, c);
                size_t o = ((((size_t)l * N + zi) * N + yi) * N + xi) * 3;
                table[o] = (float)c[0]; table[o + 1] = (float)c[1]; table[o + 2] = (float)c[2];
                worst[tid] = std::max(worst[tid], e);
            };
            for (int zi = start; zi < N; ++zi) cell(zi);
            c[0] = c[1] = c[2] = 0;
            for (int zi = start; zi >= 0; --zi) cell(zi);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    double w = 0;
    for (double x : worst) w = std::max(w, x);
    FILE* f = std::fopen(argv[2], "wb");
    if (!f) { std::perror(argv[2]); return 1; }
    std::fwrite(znodes.data(), 4, N, f);
    std::fwrite(table.data(), 4, table.size(), f);
    std::fclose(f);
    std::printf("rgb2spec_fit: wrote %s (%zu B), worst cell dE = %.4f\n", argv[2], (size_t)(N * 4 + table.size() * 4), std::sqrt(w));
    return 0;
}
