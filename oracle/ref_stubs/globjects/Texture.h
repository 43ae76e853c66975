// ORACLE BUILD STUB (test infrastructure): calibration_inverter.hpp includes this header but uses nothing from it.
#ifndef RR_REF_STUB_Texture_H
#define RR_REF_STUB_Texture_H
namespace globjects { class Texture; }
#endif
