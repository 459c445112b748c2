/* placeholder: encoder restatement lands in a later commit */
#include "jpeg_oracle.h"
