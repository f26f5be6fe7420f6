// kernel instances for complex length 2^8 (one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e8() { return build_entries<8>(); } } }
