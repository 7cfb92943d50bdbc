/* xsTypes.h shim — the one XSPEC typedef the reference needs (src/XspecSpectrum.h:22,26). */
#ifndef ORACLE_SHIM_XSTYPES_H_
#define ORACLE_SHIM_XSTYPES_H_
#include <valarray>
typedef std::valarray<double> RealArray;
#endif
