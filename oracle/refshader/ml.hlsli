// TEST INFRASTRUCTURE.  Stand-in for MathLib's ml.hlsli (an external dependency of the reference that is not vendored in
// /root/reference): every function the shaders call forwards to the oracle's restatement in oracle/mathlib.h, so that the reference
// shaders compiled through oracle/refshader/hlsl_cpp.h and the oracle's passes share one MathLib.  (This text goes through the C
// preprocessor together with the shader and is compiled as C++ -- see oracle/build_refshaders.py.)
#define ML_SPECULAR_DOMINANT_DIRECTION_G1 1
#define ML_SPECULAR_DOMINANT_DIRECTION_G2 0
#define ML_SPECULAR_DOMINANT_DIRECTION_APPROX 2
#define ML_MAP1(ns, fn) \
    inline float2 fn(float2 a) { return float2(fn(a.x), fn(a.y)); } \
    inline float3 fn(float3 a) { return float3(fn(a.x), fn(a.y), fn(a.z)); } \
    inline float4 fn(float4 a) { return float4(fn(a.x), fn(a.y), fn(a.z), fn(a.w)); }
namespace Math
{
inline float Pi(float x) { return hlsl::Math::Pi(x); }
inline float DegToRad(float x) { return hlsl::Math::DegToRad(x); }
inline float LinearStep(float a, float b, float x) { return hlsl::Math::LinearStep(a, b, x); }
inline float2 LinearStep(float2 a, float2 b, float2 x) { return float2(LinearStep(a.x, b.x, x.x), LinearStep(a.y, b.y, x.y)); }
inline float3 LinearStep(float3 a, float3 b, float3 x) { return float3(LinearStep(a.x, b.x, x.x), LinearStep(a.y, b.y, x.y), LinearStep(a.z, b.z, x.z)); }
inline float4 LinearStep(float4 a, float4 b, float4 x) { return float4(LinearStep(a.x, b.x, x.x), LinearStep(a.y, b.y, x.y), LinearStep(a.z, b.z, x.z), LinearStep(a.w, b.w, x.w)); }
inline float SmoothStep01(float x) { return hlsl::Math::SmoothStep01(x); }
ML_MAP1(Math, SmoothStep01)
inline float SmoothStep(float a, float b, float x) { return hlsl::Math::SmoothStep(a, b, x); }
inline float2 SmoothStep(float2 a, float2 b, float2 x) { return float2(SmoothStep(a.x, b.x, x.x), SmoothStep(a.y, b.y, x.y)); }
inline float3 SmoothStep(float3 a, float3 b, float3 x) { return float3(SmoothStep(a.x, b.x, x.x), SmoothStep(a.y, b.y, x.y), SmoothStep(a.z, b.z, x.z)); }
inline float4 SmoothStep(float4 a, float4 b, float4 x) { return float4(SmoothStep(a.x, b.x, x.x), SmoothStep(a.y, b.y, x.y), SmoothStep(a.z, b.z, x.z), SmoothStep(a.w, b.w, x.w)); }
inline float Sqrt01(float x) { return hlsl::Math::Sqrt01(x); }
ML_MAP1(Math, Sqrt01)
inline float Pow01(float x, float y) { return hlsl::Math::Pow01(x, y); }
inline float2 Pow01(float2 x, float y) { return float2(Pow01(x.x, y), Pow01(x.y, y)); }
inline float3 Pow01(float3 x, float y) { return float3(Pow01(x.x, y), Pow01(x.y, y), Pow01(x.z, y)); }
inline float4 Pow01(float4 x, float y) { return float4(Pow01(x.x, y), Pow01(x.y, y), Pow01(x.z, y), Pow01(x.w, y)); }
inline float Rsqrt(float x) { return hlsl::Math::Rsqrt(x); }
inline float LengthSquared(float2 v) { return hlsl::Math::LengthSquared(O(v)); }
inline float LengthSquared(float3 v) { return hlsl::Math::LengthSquared(O(v)); }
inline float PositiveRcp(float x) { return hlsl::Math::PositiveRcp(x); }
ML_MAP1(Math, PositiveRcp)
inline float AcosApprox(float x) { return hlsl::Math::AcosApprox(x); }
} // namespace Math

namespace Geometry
{
inline float4 GetRotator(float a) { return S(hlsl::Geometry::GetRotator(a)); }
inline float2 RotateVector(float4 r, float2 v) { return S(hlsl::Geometry::RotateVector(O(r), O(v))); }
inline float4 CombineRotators(float4 r1, float4 r2) { return S(hlsl::Geometry::CombineRotators(O(r1), O(r2))); }
inline float4 ScaleRotator(float4 r, float2 s) { return S(hlsl::Geometry::ScaleRotator(O(r), O(s))); }
inline float4 ScaleRotator(float4 r, float s) { return S(hlsl::Geometry::ScaleRotator(O(r), hlsl::float2(s))); }
inline float3 RotateVector(float4x4 m, float3 v) { return S(hlsl::Geometry::RotateVector(O(m), O(v))); }
inline float3 RotateVectorInverse(float4x4 m, float3 v) { return S(hlsl::Geometry::RotateVectorInverse(O(m), O(v))); }
inline float3 RotateVector(float3x3 m, float3 v) { return S(hlsl::Geometry::RotateVector(O(m), O(v))); }
inline float3 AffineTransform(float4x4 m, float3 p) { return S(hlsl::Geometry::AffineTransform(O(m), O(p))); }
inline float4 ProjectiveTransform(float4x4 m, float3 p) { return S(hlsl::Geometry::ProjectiveTransform(O(m), O(p))); }
inline float4 ProjectiveTransform(float4x4 m, float4 p) { return mul(m, p); }
inline float2 GetScreenUv(float4x4 worldToClip, float3 X, bool killBackprojection = true) { return S(hlsl::Geometry::GetScreenUv(O(worldToClip), O(X), killBackprojection)); }
inline float3 ReconstructViewPosition(float2 uv, float4 frustum, float viewZ = 1.0f, float orthoMode = 0.0f) { return S(hlsl::Geometry::ReconstructViewPosition(O(uv), O(frustum), viewZ, orthoMode)); }
inline float3x3 GetBasis(float3 N) { return S(hlsl::Geometry::GetBasis(O(N))); }
} // namespace Geometry

namespace Filtering
{
struct Bilinear { float2 origin; float2 weights; };
struct CatmullRom { float2 origin; };
inline hlsl::Filtering::Bilinear O(Bilinear f) { hlsl::Filtering::Bilinear r; r.origin = refshader::O(f.origin); r.weights = refshader::O(f.weights); return r; }
inline Bilinear GetBilinearFilter(float2 uv, float2 texSize)
{
    hlsl::Filtering::Bilinear f = hlsl::Filtering::GetBilinearFilter(refshader::O(uv), refshader::O(texSize));
    Bilinear r; r.origin = S(f.origin); r.weights = S(f.weights); return r;
}
inline float ApplyBilinearFilter(float s00, float s10, float s01, float s11, Bilinear f) { return hlsl::Filtering::ApplyBilinearFilter(s00, s10, s01, s11, O(f)); }
inline float2 ApplyBilinearFilter(float2 s00, float2 s10, float2 s01, float2 s11, Bilinear f) { return lerp(lerp(s00, s10, f.weights.x), lerp(s01, s11, f.weights.x), f.weights.y); }
inline float3 ApplyBilinearFilter(float3 s00, float3 s10, float3 s01, float3 s11, Bilinear f) { return lerp(lerp(s00, s10, f.weights.x), lerp(s01, s11, f.weights.x), f.weights.y); }
inline float4 ApplyBilinearFilter(float4 s00, float4 s10, float4 s01, float4 s11, Bilinear f) { return lerp(lerp(s00, s10, f.weights.x), lerp(s01, s11, f.weights.x), f.weights.y); }
inline float4 GetBilinearCustomWeights(Bilinear f, float4 customWeights) { return S(hlsl::Filtering::GetBilinearCustomWeights(O(f), refshader::O(customWeights))); }
inline float ApplyBilinearCustomWeights(float s00, float s10, float s01, float s11, float4 w) { return hlsl::Filtering::ApplyBilinearCustomWeights(s00, s10, s01, s11, refshader::O(w)); }
#define ML_CUSTOM_WEIGHTS(T) \
    inline T ApplyBilinearCustomWeights(T s00, T s10, T s01, T s11, float4 w) \
    { \
        float sum = dot(w, float4(1.0f)); \
        T r = s00 * w.x + s10 * w.y + s01 * w.z + s11 * w.w; \
        return sum < 0.0001f ? T(0.0f) : r / sum; \
    }
ML_CUSTOM_WEIGHTS(float2) ML_CUSTOM_WEIGHTS(float3) ML_CUSTOM_WEIGHTS(float4)
inline CatmullRom GetCatmullRomFilter(float2 uv, float2 texSize) { CatmullRom r; r.origin = S(hlsl::Filtering::GetCatmullRomFilter(refshader::O(uv), refshader::O(texSize)).origin); return r; }
inline float GetModifiedRoughnessFromNormalVariance(float linearRoughness, float3 nonNormalizedAverageNormal) { return hlsl::Filtering::GetModifiedRoughnessFromNormalVariance(linearRoughness, refshader::O(nonNormalizedAverageNormal)); }
} // namespace Filtering

namespace Packing
{
inline uint RgbaToUint(float4 c, uint rb, uint gb, uint bb, uint ab) { return hlsl::Packing::RgbaToUint(O(c), rb, gb, bb, ab); }
inline float4 UintToRgba(uint p, uint rb, uint gb, uint bb, uint ab) { return S(hlsl::Packing::UintToRgba(p, rb, gb, bb, ab)); }
} // namespace Packing

namespace ImportanceSampling
{
inline float GetSpecularDominantFactor(float NoV, float roughness, int = 0 /* ML_SPECULAR_DOMINANT_DIRECTION_G2 */) { return hlsl::ImportanceSampling::GetSpecularDominantFactor(NoV, roughness); }
inline float4 GetSpecularDominantDirection(float3 N, float3 V, float roughness, int = 0 /* ML_SPECULAR_DOMINANT_DIRECTION_G2: the fit the oracle restates */) { return S(hlsl::ImportanceSampling::GetSpecularDominantDirection(O(N), O(V), roughness)); }
inline float GetSpecularLobeTanHalfAngle(float roughness, float percentOfVolume = 0.75f) { return hlsl::ImportanceSampling::GetSpecularLobeTanHalfAngle(roughness, percentOfVolume); }
} // namespace ImportanceSampling

namespace Color
{
inline float Luminance(float3 c) { return hlsl::Color::Luminance(O(c)); }
// (restated like oracle/relax.cpp RgbToYCoCg / YCoCgToRgb)
inline float3 RgbToYCoCg(float3 c) { return float3(dot(c, float3(0.25f, 0.5f, 0.25f)), dot(c, float3(0.5f, 0.0f, -0.5f)), dot(c, float3(-0.25f, 0.5f, -0.25f))); }
inline float3 YCoCgToRgb(float3 c)
{
    float t = c.x - c.z;
    float3 r;
    r.y = c.x + c.z;
    r.x = t + c.y;
    r.z = t - c.y;
    return max(r, float3(0.0f));
}
inline float Clamp(float m1, float sigma, float x) { return hlsl::Color::Clamp(m1, sigma, x); }
inline float2 Clamp(float2 m1, float2 sigma, float2 x) { return clamp(x, m1 - sigma, m1 + sigma); }
inline float3 Clamp(float3 m1, float3 sigma, float3 x) { return clamp(x, m1 - sigma, m1 + sigma); }
inline float4 Clamp(float4 m1, float4 sigma, float4 x) { return clamp(x, m1 - sigma, m1 + sigma); }
} // namespace Color

namespace BRDF
{
inline float Pow5(float x) { return hlsl::BRDF::Pow5(x); }
inline void ConvertBaseColorMetalnessToAlbedoRf0(float3 baseColor, float metalness, float3& albedo, float3& Rf0)
{
    hlsl::float3 a, r;
    hlsl::BRDF::ConvertBaseColorMetalnessToAlbedoRf0(O(baseColor), metalness, a, r);
    albedo = S(a);
    Rf0 = S(r);
}
inline float3 EnvironmentTerm_Rtg(float3 Rf0, float NoV, float linearRoughness) { return S(hlsl::BRDF::EnvironmentTerm_Rtg(O(Rf0), NoV, linearRoughness)); }
} // namespace BRDF

namespace Sequence
{
inline uint CheckerBoard(int2 p, uint frameIndex) { return hlsl::Sequence::CheckerBoard(O(p), frameIndex); }
inline uint CheckerBoard(uint2 p, uint frameIndex) { return hlsl::Sequence::CheckerBoard(hlsl::int2((int)p.x, (int)p.y), frameIndex); }
inline float Bayer4x4(int2 p, uint frameIndex) { return hlsl::Sequence::Bayer4x4(O(p), frameIndex); }
inline float Bayer4x4(uint2 p, uint frameIndex) { return hlsl::Sequence::Bayer4x4(hlsl::int2((int)p.x, (int)p.y), frameIndex); }
} // namespace Sequence

// Rng::Hash keeps its seed in a per-thread global of the shader; here: one state per fiber of the thread group
namespace Rng
{
namespace Hash
{
inline hlsl::RngHash& State()
{
    static hlsl::RngHash states[RefShaderGroup::kMaxThreads];
    return states[RefShaderCurrentGroup().current];
}
inline void Initialize(int2 pixelPos, uint frameIndex) { State().Initialize(O(pixelPos), frameIndex); }
inline void Initialize(uint2 pixelPos, uint frameIndex) { State().Initialize(hlsl::int2((int)pixelPos.x, (int)pixelPos.y), frameIndex); }
inline float GetFloat() { return State().GetFloat(); }
inline float2 GetFloat2() { return S(State().GetFloat2()); }
inline float4 GetFloat4() { float2 a = GetFloat2(); float2 b = GetFloat2(); return float4(a, b); }
} // namespace Hash
} // namespace Rng
