/* esl_scorematrix.h -- Easel compat shim: everything lives in easel.h (see that file). */
#include "easel.h"
