/*
 * StateValidityCheckerRSS.h -- drop-in for validity_checkers/StateValidityCheckerRSS{,_PRM}.{h,cpp}: the PCS class plus
 * random singular sampling (sampleSingular) and the RBS variant with a singular endpoint.
 * Compile StateValidityCheckerPCS.cpp with -DCKC_FLAVOUR_RSS.
 */
#ifndef CKC_FLAVOUR_RSS
#define CKC_FLAVOUR_RSS
#endif
#include "StateValidityCheckerPCS.h"
