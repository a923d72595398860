// capi.cu — version and status text of the C ABI (include/cloudaae_b200.h).
#include "common.cuh"

extern "C" int caae_abi_version(void) { return CAAE_ABI_VERSION; }

extern "C" const char* caae_status_string(int status) {
  switch (status) {
    case CAAE_OK: return "ok";
    case CAAE_E_BADSHAPE: return "invalid size argument";
    case CAAE_E_NULLPTR: return "null pointer for a non-empty tensor";
    case CAAE_E_SCRATCH: return "scratch buffer required for this size but not provided";
    case CAAE_E_UNSUPPORTED: return "unsupported configuration";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "unknown cloudaae_b200 status";
}

// CRC-32C (Castagnoli) of a host buffer, continuing from `crc` (0 to start): the per-tensor checksum of
// TensorFlow's checkpoint format (cloudaae_b200/data/tf_checkpoint.py).  Host code, slicing-by-8.
extern "C" unsigned int caae_crc32c(unsigned int crc, const void* data, unsigned long long n) {
  static unsigned int tbl[8][256];
  static bool ready = false;
  if (!ready) {
    for (unsigned int i = 0; i < 256; ++i) {
      unsigned int c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      tbl[0][i] = c;
    }
    for (unsigned int i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) tbl[t][i] = (tbl[t - 1][i] >> 8) ^ tbl[0][tbl[t - 1][i] & 0xFFu];
    ready = true;
  }
  const unsigned char* p = static_cast<const unsigned char*>(data);
  unsigned int c = ~crc;
  while (n >= 8) {
    const unsigned int lo = c ^ ((unsigned int)p[0] | ((unsigned int)p[1] << 8) | ((unsigned int)p[2] << 16) | ((unsigned int)p[3] << 24));
    c = tbl[7][lo & 0xFFu] ^ tbl[6][(lo >> 8) & 0xFFu] ^ tbl[5][(lo >> 16) & 0xFFu] ^ tbl[4][lo >> 24] ^
        tbl[3][p[4]] ^ tbl[2][p[5]] ^ tbl[1][p[6]] ^ tbl[0][p[7]];
    p += 8; n -= 8;
  }
  while (n--) c = tbl[0][(c ^ *p++) & 0xFFu] ^ (c >> 8);
  return ~c;
}

#include <stdlib.h>
namespace caae {
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CAAE_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace caae
