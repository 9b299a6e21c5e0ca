/* TEST INFRASTRUCTURE.  The resource-declaration macros of NRD.hlsli's "custom engine" branch (NRD.hlsli:96-114), defined so that a
   shader's constants, inputs and outputs become C++ globals that register themselves with the driver in oracle/refshader/hlsl_cpp.h.
   Seen by the C preprocessor only (oracle/build_refshaders.py). */
#define NRD_CONSTANTS_START( resourceName )
#define NRD_CONSTANT( constantType, constantName ) constantType constantName; static RefShaderReg _rs_##constantName( &constantName, sizeof( constantType ), #constantName, 0 );
#define NRD_CONSTANTS_END
#define NRD_INPUTS_START
#define NRD_INPUT( resourceType, resourceName, regName, bindingIndex ) resourceType resourceName; static RefShaderReg _rs_##resourceName( RefShaderTexPtr( &resourceName ), 0, #resourceName, 1 );
#define NRD_INPUTS_END
#define NRD_OUTPUTS_START
#define NRD_OUTPUT( resourceType, resourceName, regName, bindingIndex ) resourceType resourceName; static RefShaderReg _rs_##resourceName( RefShaderTexPtr( &resourceName ), 0, #resourceName, 2 );
#define NRD_OUTPUTS_END
#define NRD_SAMPLERS_START
#define NRD_SAMPLER( resourceType, resourceName, regName, bindingIndex ) resourceType resourceName = { bindingIndex };
#define NRD_SAMPLERS_END
#define NRD_CS_MAIN refshader_main
#define compiletime
