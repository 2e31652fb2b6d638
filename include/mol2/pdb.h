/* libmol2 header name kept so reference-style includes resolve; see mini.h */
#include "mini.h"
