// kernel instances for complex length 2^13 (C2C only; one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e13() { return build_entries_large<13>(); } } }
