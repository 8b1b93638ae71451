#include "shim/prelude.h"
