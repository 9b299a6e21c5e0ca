// Host-side matrix / sequence helpers for the pass scheduler.
//
// The reference takes these from MathLib's ml.h (NVIDIA-RTX/MathLib, fetched at configure time with GIT_TAG main,
// reference CMakeLists.txt:120-129) which is NOT present in /root/reference.  They are restated here from their
// published behaviour and from how the reference calls them (Source/InstanceImpl.cpp:339-456, Source/Relax.cpp:53-78,
// Source/Sigma.cpp:107).  Conventions: column-major storage, column vectors, clip = M * v.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace nrdb200
{
struct Vec3 { float x, y, z; };
struct Vec4 { float x, y, z, w; };

struct Mat4
{
    // c[i] is column i; element (row r, col c) = m[c * 4 + r]
    float m[16];

    static Mat4 Identity()
    {
        Mat4 r{};
        r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
        return r;
    }
    static Mat4 FromColumnMajor(const float* p)
    {
        Mat4 r;
        memcpy(r.m, p, sizeof(r.m));
        return r;
    }
    float& at(int row, int col) { return m[col * 4 + row]; }
    float at(int row, int col) const { return m[col * 4 + row]; }

    Mat4 operator*(const Mat4& b) const
    {
        Mat4 r;
        for (int c = 0; c < 4; c++)
            for (int row = 0; row < 4; row++)
            {
                float s = 0.0f;
                for (int k = 0; k < 4; k++)
                    s += at(row, k) * b.at(k, c);
                r.at(row, c) = s;
            }
        return r;
    }
    Vec4 mul(const Vec4& v) const
    {
        return {at(0, 0) * v.x + at(0, 1) * v.y + at(0, 2) * v.z + at(0, 3) * v.w,
                at(1, 0) * v.x + at(1, 1) * v.y + at(1, 2) * v.z + at(1, 3) * v.w,
                at(2, 0) * v.x + at(2, 1) * v.y + at(2, 2) * v.z + at(2, 3) * v.w,
                at(3, 0) * v.x + at(3, 1) * v.y + at(3, 2) * v.z + at(3, 3) * v.w};
    }
    void negateColumn(int c)
    {
        for (int r = 0; r < 4; r++) at(r, c) = -at(r, c);
    }
    void negateRow(int r)
    {
        for (int c = 0; c < 4; c++) at(r, c) = -at(r, c);
    }
    Vec3 translation() const { return {at(0, 3), at(1, 3), at(2, 3)}; }
    void setTranslation(const Vec3& t)
    {
        at(0, 3) = t.x;
        at(1, 3) = t.y;
        at(2, 3) = t.z;
    }
    // inverse of a rigid transform [R | t]: [R^T | -R^T t]
    void invertOrtho()
    {
        Mat4 r = Identity();
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                r.at(i, j) = at(j, i);
        Vec3 t = translation();
        r.at(0, 3) = -(r.at(0, 0) * t.x + r.at(0, 1) * t.y + r.at(0, 2) * t.z);
        r.at(1, 3) = -(r.at(1, 0) * t.x + r.at(1, 1) * t.y + r.at(1, 2) * t.z);
        r.at(2, 3) = -(r.at(2, 0) * t.x + r.at(2, 1) * t.y + r.at(2, 2) * t.z);
        *this = r;
    }
    // general inverse (Gauss-Jordan with partial pivoting, evaluated in double)
    void invert()
    {
        double a[4][8];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++)
            {
                a[r][c] = at(r, c);
                a[r][c + 4] = r == c ? 1.0 : 0.0;
            }
        for (int col = 0; col < 4; col++)
        {
            int piv = col;
            for (int r = col + 1; r < 4; r++)
                if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
            if (a[piv][col] == 0.0) { *this = Identity(); return; }
            if (piv != col)
                for (int c = 0; c < 8; c++) { double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
            double inv = 1.0 / a[col][col];
            for (int c = 0; c < 8; c++) a[col][c] *= inv;
            for (int r = 0; r < 4; r++)
                if (r != col)
                {
                    double f = a[r][col];
                    if (f != 0.0)
                        for (int c = 0; c < 8; c++) a[r][c] -= f * a[col][c];
                }
        }
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++)
                at(r, c) = (float)a[r][c + 4];
    }
};

enum ProjectionFlags : uint32_t { PROJ_ORTHO = 1, PROJ_LEFT_HANDED = 2 };

// Restates what the scheduler needs from ml.h DecomposeProjection (call sites InstanceImpl.cpp:394,446,451):
//  - flags: left-handed iff clip.w grows with +z (perspective) / depth grows with +z (ortho); ortho iff last row = (0,0,0,1)
//  - frustum (x0, y0, dx, dy) such that viewPos.xy = (uv * frustum.zw + frustum.xy) * viewZ reproduces the projection
//    (uv.y grows downwards), i.e. the inverse of uv = (clip.xy / clip.w) * (0.5, -0.5) + 0.5
//  - project[1] = y scale of the projection (cot(fovY / 2) for perspective)
// The frustum is derived for the matrix as given; the scheduler negates the z column first when it is right-handed.
inline void DecomposeProjection(const Mat4& p, uint32_t& flags, float frustum[4], float project[3])
{
    flags = 0;
    const bool isOrtho = p.at(3, 0) == 0.0f && p.at(3, 1) == 0.0f && p.at(3, 2) == 0.0f && p.at(3, 3) == 1.0f;
    const float m00 = p.at(0, 0), m11 = p.at(1, 1);
    if (isOrtho)
    {
        flags |= PROJ_ORTHO;
        if (p.at(2, 2) > 0.0f) flags |= PROJ_LEFT_HANDED;
        // x_ndc = m00 * x + m03  =>  x = (2u - 1 - m03) / m00 ; ReconstructViewPosition multiplies by orthoMode = -1
        const float m03 = p.at(0, 3), m13 = p.at(1, 3);
        frustum[0] = -(-1.0f - m03) / m00;
        frustum[1] = -(1.0f - m13) / m11;
        frustum[2] = -2.0f / m00;
        frustum[3] = 2.0f / m11;
    }
    else
    {
        const float wz = p.at(3, 2); // clip.w = wz * z
        if (wz > 0.0f) flags |= PROJ_LEFT_HANDED;
        const float s = wz > 0.0f ? 1.0f : -1.0f;
        const float m02 = p.at(0, 2) * s, m12 = p.at(1, 2) * s;
        frustum[0] = (-1.0f - m02) / m00;
        frustum[1] = (1.0f - m12) / m11;
        frustum[2] = 2.0f / m00;
        frustum[3] = -2.0f / m11;
    }
    if (project)
    {
        project[0] = m00;
        project[1] = m11;
        project[2] = p.at(2, 2);
    }
}

inline float Radians(float deg) { return deg * 3.14159265358979323846f / 180.0f; }

// Sequence::Weyl1D(p, n) = frac(p + n * phi^-1) with the golden-ratio step held as a 24-bit fixed-point integer
inline float Weyl1D(float p, uint32_t n)
{
    float v = p + float(n * 10368889u) / 16777216.0f;
    return v - std::floor(v);
}

// Sequence::Bayer4x4(pixel, frame): 4x4 ordered-dither matrix value in [0,1), advanced by the frame index
inline uint32_t Bayer4x4ui(uint32_t x, uint32_t y, uint32_t frameIndex)
{
    static const uint32_t k[4][4] = {{0, 8, 2, 10}, {12, 4, 14, 6}, {3, 11, 1, 9}, {15, 7, 13, 5}};
    return (k[y & 3][x & 3] + frameIndex) & 0xF;
}
inline float Bayer4x4(uint32_t x, uint32_t y, uint32_t frameIndex) { return float(Bayer4x4ui(x, y, frameIndex)) / 16.0f; }

// Geometry::GetRotator(angle) = (cos, sin, -sin, cos); RotateVector(r, v) = v.x * r.xz + v.y * r.yw
inline Vec4 GetRotator(float angle)
{
    float ca = std::cos(angle), sa = std::sin(angle);
    return {ca, sa, -sa, ca};
}
// 2x2 product of two rotators
inline Vec4 CombineRotators(const Vec4& r1, const Vec4& r2)
{
    return {r1.x * r2.x + r1.z * r2.y, r1.y * r2.x + r1.w * r2.y, r1.x * r2.z + r1.z * r2.w, r1.y * r2.z + r1.w * r2.w};
}
} // namespace nrdb200
