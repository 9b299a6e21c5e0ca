// Per-dispatch constant blocks.  The byte layout is part of the DispatchDesc contract
// (DispatchDesc::constantBufferData): it is the HLSL cbuffer packing of the reference's
//   REBLUR_SHARED_CONSTANTS  Shaders/Include/REBLUR_Config.hlsli:113-186   (832 B)
//   RELAX_SHARED_CONSTANTS   Shaders/Include/RELAX_Config.hlsli:21-99      (+ gStepSize, gIsLastPass for A-trous,
//                            Shaders/Resources/RELAX_Atrous.resources.hlsli:11-15)
//   SIGMA_SHARED_CONSTANTS   Shaders/Include/SIGMA_Config.hlsli:46-80      (528 B)
// float4x4 = 4 consecutive float4 columns (column_major), no member straddles a 16-byte register.
// Shared by the host scheduler (fills them) and the CUDA kernels (receive them as __grid_constant__).
#pragma once
#include <cstdint>

namespace nrdb200
{
struct alignas(16) ReblurConstants
{
    float gWorldToClip[16];
    float gViewToClip[16];
    float gViewToWorld[16];
    float gWorldToViewPrev[16];
    float gWorldToClipPrev[16];
    float gWorldPrevToWorld[16];
    float gRotatorPre[4];
    float gRotator[4];
    float gRotatorPost[4];
    float gFrustum[4];
    float gFrustumPrev[4];
    float gCameraDelta[4];
    float gHitDistParams[4];
    float gViewVectorWorld[4];
    float gViewVectorWorldPrev[4];
    float gMvScale[4];
    float gAntilagParams[2];
    float gResourceSize[2];
    float gResourceSizeInv[2];
    float gResourceSizeInvPrev[2];
    float gRectSize[2];
    float gRectSizeInv[2];
    float gRectSizePrev[2];
    float gResolutionScale[2];
    float gResolutionScalePrev[2];
    float gRectOffset[2];
    float gSpecProbabilityThresholdsForMvModification[2];
    float gJitter[2];
    uint32_t gPrintfAt[2];
    uint32_t gRectOrigin[2];
    int32_t gRectSizeMinusOne[2];
    float gDisocclusionThreshold;
    float gDisocclusionThresholdAlternate;
    float gCameraAttachedReflectionMaterialID;
    float gStrandMaterialID;
    float gStrandThickness;
    float gStabilizationStrength;
    float gHitDistStabilizationStrength;
    float gDebug;
    float gOrthoMode;
    float gUnproject;
    float gDenoisingRange;
    float gPlaneDistSensitivity;
    float gFramerateScale;
    float gMinBlurRadius;
    float gMaxBlurRadius;
    float gDiffPrepassBlurRadius;
    float gSpecPrepassBlurRadius;
    float gMaxAccumulatedFrameNum;
    float gMaxFastAccumulatedFrameNum;
    float gAntiFirefly;
    float gLobeAngleFraction;
    float gRoughnessFraction;
    float gResponsiveAccumulationRoughnessThreshold;
    float gHistoryFixFrameNum;
    float gHistoryFixBasePixelStride;
    float gMinRectDimMulUnproject;
    float gUsePrepassNotOnlyForSpecularMotionEstimation;
    float gSplitScreen;
    float gSplitScreenPrev;
    float gCheckerboardResolveAccumSpeed;
    float gViewZScale;
    float gFireflySuppressorMinRelativeScale;
    float gMinHitDistanceWeight;
    float gDiffMinMaterial;
    float gSpecMinMaterial;
    uint32_t gHasHistoryConfidence;
    uint32_t gHasDisocclusionThresholdMix;
    uint32_t gDiffCheckerboard;
    uint32_t gSpecCheckerboard;
    uint32_t gFrameIndex;
    uint32_t gIsRectChanged;
    uint32_t gResetHistory;
};
static_assert(sizeof(ReblurConstants) == 832, "REBLUR constant block must be 832 bytes");

struct alignas(16) RelaxConstants
{
    float gWorldToClip[16];
    float gWorldToClipPrev[16];
    float gWorldToViewPrev[16];
    float gWorldPrevToWorld[16];
    float gRotatorPre[4];
    float gFrustumRight[4];
    float gFrustumUp[4];
    float gFrustumForward[4];
    float gPrevFrustumRight[4];
    float gPrevFrustumUp[4];
    float gPrevFrustumForward[4];
    float gCameraDelta[4];
    float gMvScale[4];
    float gJitter[2];
    float gResolutionScale[2];
    float gRectOffset[2];
    float gResourceSizeInv[2];
    float gResourceSize[2];
    float gRectSizeInv[2];
    float gRectSizePrev[2];
    float gResourceSizeInvPrev[2];
    uint32_t gPrintfAt[2];
    uint32_t gRectOrigin[2];
    int32_t gRectSize[2];
    float gSpecMaxAccumulatedFrameNum;
    float gSpecMaxFastAccumulatedFrameNum;
    float gDiffMaxAccumulatedFrameNum;
    float gDiffMaxFastAccumulatedFrameNum;
    float gDisocclusionThreshold;
    float gDisocclusionThresholdAlternate;
    float gCameraAttachedReflectionMaterialID;
    float gStrandMaterialID;
    float gStrandThickness;
    float gRoughnessFraction;
    float gSpecVarianceBoost;
    float gSplitScreen;
    float gDiffBlurRadius;
    float gSpecBlurRadius;
    float gDepthThreshold;
    float gLobeAngleFraction;
    float gSpecLobeAngleSlack;
    float gHistoryFixEdgeStoppingNormalPower;
    float gRoughnessEdgeStoppingRelaxation;
    float gNormalEdgeStoppingRelaxation;
    float gColorBoxSigmaScale;
    float gHistoryAccelerationAmount;
    float gHistoryResetTemporalSigmaScale;
    float gHistoryResetSpatialSigmaScale;
    float gHistoryResetAmount;
    float gDenoisingRange;
    float gSpecPhiLuminance;
    float gDiffPhiLuminance;
    float gDiffMaxLuminanceRelativeDifference;
    float gSpecMaxLuminanceRelativeDifference;
    float gLuminanceEdgeStoppingRelaxation;
    float gConfidenceDrivenRelaxationMultiplier;
    float gConfidenceDrivenLuminanceEdgeStoppingRelaxation;
    float gConfidenceDrivenNormalEdgeStoppingRelaxation;
    float gDebug;
    float gOrthoMode;
    float gUnproject;
    float gFramerateScale;
    float gCheckerboardResolveAccumSpeed;
    float gJitterDelta;
    float gHistoryFixFrameNum;
    float gHistoryFixBasePixelStride;
    float gHistoryThreshold;
    float gViewZScale;
    float gMinHitDistanceWeight;
    float gDiffMinMaterial;
    float gSpecMinMaterial;
    uint32_t gRoughnessEdgeStoppingEnabled;
    uint32_t gFrameIndex;
    uint32_t gDiffCheckerboard;
    uint32_t gSpecCheckerboard;
    uint32_t gHasHistoryConfidence;
    uint32_t gHasDisocclusionThresholdMix;
    uint32_t gResetHistory;
    // present in every RELAX block; only the A-trous dispatches give them a meaning
    uint32_t gStepSize;
    uint32_t gIsLastPass;
};
static_assert(sizeof(RelaxConstants) % 16 == 0, "RELAX constant block must be register aligned");

struct alignas(16) SigmaConstants
{
    float gWorldToView[16];
    float gViewToClip[16];
    float gWorldToClipPrev[16];
    float gWorldToViewPrev[16];
    float gRotator[4];
    float gRotatorPost[4];
    float gViewVectorWorld[4];
    float gLightDirectionView[4];
    float gFrustum[4];
    float gFrustumPrev[4];
    float gCameraDelta[4];
    float gMvScale[4];
    float gResourceSizeInv[2];
    float gResourceSizeInvPrev[2];
    float gRectSize[2];
    float gRectSizeInv[2];
    float gRectSizePrev[2];
    float gResolutionScale[2];
    float gRectOffset[2];
    uint32_t gPrintfAt[2];
    uint32_t gRectOrigin[2];
    int32_t gRectSizeMinusOne[2];
    int32_t gTilesSizeMinusOne[2];
    float gOrthoMode;
    float gUnproject;
    float gDenoisingRange;
    float gPlaneDistSensitivity;
    float gStabilizationStrength;
    float gDebug;
    float gSplitScreen;
    float gViewZScale;
    float gMinRectDimMulUnproject;
    uint32_t gFrameIndex;
    uint32_t gIsRectChanged;
};
static_assert(sizeof(SigmaConstants) == 528, "SIGMA constant block must be 528 bytes");

// REFERENCE_TemporalAccumulation.resources.hlsli:11-16 / REFERENCE_Copy.resources.hlsli:11-16 (20 bytes each in the reference)
struct ReferenceAccumulateConstants
{
    uint32_t gRectOrigin[2];
    float gAccumSpeed;
    float gDebug;
    float gViewZScale;
};
struct ReferenceCopyConstants
{
    float gRectSizeInv[2];
    float gSplitScreen;
    float gDebug;
    float gViewZScale;
};
static_assert(sizeof(ReferenceAccumulateConstants) == 20 && sizeof(ReferenceCopyConstants) == 20, "REFERENCE constant blocks are 20 bytes");
} // namespace nrdb200
