/* StateValidityCheckerRX.h -- drop-in for validity_checkers/StateValidityCheckerRX.h: the relaxed-constraint checker shares
 * the GD drop-in's class (kdl chain FK + collision), selected by CKC_FLAVOUR_RX. */
#ifndef CKC_FLAVOUR_RX
#define CKC_FLAVOUR_RX
#endif
#include "StateValidityCheckerGD.h"
