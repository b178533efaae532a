"""fg_idprimer_b200: B200-native (sm_100a) IdentifyPrimers matching path behind the C-ABI in include/idprimer.h."""
from ._lib import (F_FIRST, F_FR_PAIR, F_MATE_NEXT, F_PAIRED, F_REVERSE, F_SECOND, F_UNMAPPED, GAPPED, LOCATION, MODE_GLOBAL,
                   MODE_GLOCAL, MODE_LOCAL, NEXT_NA, NONE, RESULT_DTYPE, UNGAPPED, IdpError, lib)
from .api import (AsyncCudaAligner, GappedAlignmentPrimerMatch, IdentifyPrimers, LocationBasedPrimerMatch, Primer, ReadBatch,
                  UngappedAlignmentPrimerMatch)

__all__ = [n for n in dir() if not n.startswith("_")]
