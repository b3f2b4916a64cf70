#include "../../mental-poker_b200/csrc/transcript.hpp"
#include <chrono>
#include <vector>
#include <stdio.h>
int main() {
  size_t bytes = (17 << 20) + 37;
  std::vector<uint8_t> buf(bytes);
  for (size_t i = 0; i < bytes; i++) buf[i] = (uint8_t)(i * 131 + (i >> 8));
  for (int rep = 0; rep < 4; rep++) {
    auto t0 = std::chrono::steady_clock::now();
    mp::Blake2s h; h.update(buf.data(), 100); h.update(buf.data() + 100, buf.size() - 100); uint8_t out[32]; h.finish(out);
    auto t1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    printf("product: %.2f ms (%.0f MB/s) %02x%02x%02x%02x\n", ms, bytes / 1e3 / ms, out[0], out[1], out[2], out[31]);
  }
}
