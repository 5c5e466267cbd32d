/*
 * StateValidityCheckerHB.h -- drop-in for validity_checkers/StateValidityCheckerHB.{h,cpp}: the PCS class plus the GD
 * projection members (sampling / projection by KDL-style Newton, local connection by APC bisection).
 * Compile StateValidityCheckerPCS.cpp with -DCKC_FLAVOUR_HB.
 */
#ifndef CKC_FLAVOUR_HB
#define CKC_FLAVOUR_HB
#endif
#include "StateValidityCheckerPCS.h"
