// tangent directions per gradient pass (Dual<G3_NTAN>)
#pragma once
#ifndef G3_NTAN
#define G3_NTAN 4
#endif
