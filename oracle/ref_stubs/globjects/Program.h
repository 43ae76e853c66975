// ORACLE BUILD STUB (test infrastructure): calibration_inverter.hpp includes this header but uses nothing from it.
#ifndef RR_REF_STUB_Program_H
#define RR_REF_STUB_Program_H
namespace globjects { class Program; }
#endif
