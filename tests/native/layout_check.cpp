// Host emulation of the byte-field native layout (amq_b200/csrc/layout.cuh): packs random codes with field_src,
// builds the integer activation slots with lane_reg, runs the masked-register IMMA arithmetic of the decode kernel
// (rho_word / rho_mask) in plain C++ and checks   sum == 2^smax * sum_k code[k] * X[k]   for every row, plus the
// consistency of slot_src with both sides.  Built and run by tests/test_cpu_host.py (g++, no GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../amq_b200/csrc/layout.cuh"
using namespace amqb;

static int check_bits(int bits) {
  const int NWR = words_per_row(bits), NM = mmas_per_group(bits), smax = shift_max(bits);
  srand(1234 + bits);
  int fails = 0;
  for (int trial = 0; trial < 20; ++trial) {
    uint8_t code[16][128];
    int X[128];
    for (int r = 0; r < 16; ++r) for (int k = 0; k < 128; ++k) code[r][k] = rand() % (1 << bits);
    for (int k = 0; k < 128; ++k) X[k] = (rand() % (1 << (x_int_bits(bits) + 1))) - (1 << x_int_bits(bits)) + 1;
    // ---- pack (pack_native_codes_kernel)
    uint32_t W[32][2][4];   // [lane][row half][word]
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      for (int r = 0; r < 2; ++r)
        for (int j = 0; j < NWR; ++j) {
          uint32_t w = 0;
          for (int beta = 0; beta < 4; ++beta)
            for (int f = 0; f < n_fields(bits); ++f) {
              const FieldSrc s = field_src(bits, j, beta, f, t);
              if (s.nbits == 0) continue;
              const uint32_t q = code[g + 8 * r][s.k];
              w |= ((q >> s.lsb) & ((1u << s.nbits) - 1u)) << (s.pos + 8 * beta);
            }
          W[lane][r][j] = w;
        }
    }
    // every code bit stored exactly once: unpack and compare
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      for (int r = 0; r < 2; ++r) {
        uint8_t q[128]; int cnt[128];
        memset(q, 0, sizeof q); memset(cnt, 0, sizeof cnt);
        for (int j = 0; j < NWR; ++j)
          for (int beta = 0; beta < 4; ++beta)
            for (int f = 0; f < n_fields(bits); ++f) {
              const FieldSrc s = field_src(bits, j, beta, f, t);
              if (s.nbits == 0) continue;
              const uint32_t v = (W[lane][r][j] >> (8 * beta + s.pos)) & ((1u << s.nbits) - 1u);
              q[s.k] |= v << s.lsb; cnt[s.k] += s.nbits;
            }
        for (int i = 0; i < 16; ++i) for (int e = 0; e < 2; ++e) {
          const int k = 8 * i + 2 * t + e;
          if (cnt[k] != bits || q[k] != code[g + 8 * r][k]) { if (fails++ < 5) printf("bits %d unpack mismatch lane %d r %d k %d\n", bits, lane, r, k); }
        }
      }
    }
    // ---- activation slots (place_item / lane_reg)
    std::vector<long long> xp(NM * 32, 0);
    std::vector<int> written(NM * 32, 0);
    for (int lane = 0; lane < 32; ++lane) {
      const int I = lane >> 2, t = lane & 3;
      for (int idx = 0; idx < lane_regs(bits, I); ++idx) {
        const LaneReg R = lane_reg(bits, I, idx);
        for (int beta = 0; beta < 4; ++beta) {
          const int k = 16 * I + 8 * (beta & 1) + 2 * t + (beta >> 1);
          const int sl = R.half * 16 + 4 * t + beta;
          xp[R.m * 32 + sl] = (long long)X[k] << R.up;
          written[R.m * 32 + sl]++;
          const SlotSrc ss = slot_src(bits, R.m, sl);
          if (ss.k != k || smax - ss.shift != R.up) { if (fails++ < 5) printf("bits %d slot_src mismatch m %d sl %d: k %d vs %d, up %d vs %d\n", bits, R.m, sl, ss.k, k, smax - ss.shift, R.up); }
        }
      }
    }
    for (int i = 0; i < NM * 32; ++i) if (written[i] != 1) { if (fails++ < 5) printf("bits %d slot %d written %d times\n", bits, i, written[i]); }
    // ---- IMMA emulation (process_record): row = g + 8 r
    for (int g = 0; g < 8; ++g)
      for (int r = 0; r < 2; ++r) {
        long long sum = 0;
        for (int m = 0; m < NM; ++m)
          for (int half = 0; half < 2; ++half) {
            const int rho = 2 * m + half;
            for (int t = 0; t < 4; ++t) {
              const uint32_t a = W[g * 4 + t][r][rho_word(bits, rho)] & rho_mask(bits, rho);
              for (int beta = 0; beta < 4; ++beta) sum += (long long)((a >> (8 * beta)) & 0xFF) * xp[m * 32 + half * 16 + 4 * t + beta];
            }
          }
        long long ref = 0;
        for (int k = 0; k < 128; ++k) ref += (long long)code[g + 8 * r][k] * X[k];
        ref <<= smax;
        if (sum != ref) { if (fails++ < 5) printf("bits %d row %d: imma sum %lld != ref %lld\n", bits, g + 8 * r, sum, ref); }
      }
  }
  return fails;
}

int main() {
  int fails = 0;
  for (int bits = 2; bits <= 4; ++bits) fails += check_bits(bits);
  if (fails) { printf("FAILED: %d\n", fails); return 1; }
  printf("layout ok\n");
  return 0;
}
