// Host unit test of the safegcd Fq inversion (typlonk_b200/csrc/fq_inv.cuh): reads hex values (one
// per line, < q) on stdin, prints x^-1 mod q per line as hex, "FAIL" if the division steps did not end.
//   g++ -O2 -o fq_inv_test fq_inv_test.cpp
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>
#include "../../typlonk_b200/csrc/fq_inv.cuh"
int main() {
  std::string line;
  while (std::getline(std::cin, line)) {
    if (line.empty()) continue;
    uint32_t x[12] = {0}, out[12];
    // parse big-endian hex
    int nib = 0;
    for (int i = (int)line.size() - 1; i >= 0 && nib < 96; i--, nib++) {
      char c = line[i];
      uint32_t v = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
      x[nib / 8] |= v << (4 * (nib % 8));
    }
    bool ok = tp::fqinv_plain(x, out);
    if (!ok) { printf("FAIL\n"); continue; }
    for (int k = 11; k >= 0; k--) printf("%08x", out[k]);
    printf("\n");
  }
  return 0;
}
