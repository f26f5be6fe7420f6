// kernel instances for complex length 2^14 (C2C only; one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e14() { return build_entries_large<14>(); } } }
