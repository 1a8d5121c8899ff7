// dev tool: host primitive speeds on this box (Fr multiply, SHA-256 streams).
//   g++ -O3 -funroll-loops -std=c++17 -march=x86-64-v2 -Ibulletproofspp_b200/csrc tools/hostbench/hostbench.cpp -o /tmp/hostbench
#include <stdio.h>
#include <time.h>
#include <string>
#include <vector>
#include "host/fr64.hpp"
#include "host/sha256.hpp"
using namespace bppp;
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e9 + t.tv_nsec; }
int main() {
    h64::Fr a = h64::from_u128(12345), b = h64::from_u128(987654321);
    a = h64::mul(a, b); b = h64::sqr(a);
    uint8_t o[32];
    {
        h64::Fr x[8];
        for (int k = 0; k < 8; k++) x[k] = h64::from_u128(k + 3);
        double t0 = now();
        for (int i = 0; i < 4000000; i++) for (int k = 0; k < 8; k++) x[k] = h64::mul(x[k], b);
        double t1 = now();
        for (int k = 1; k < 8; k++) x[0] = h64::add(x[0], x[k]);
        h64::to_bytes(o, x[0]);
        printf("Fr mul (asm path if built): %.2f ns (8 independent chains) %02x\n", (t1 - t0) / 3.2e7, o[0]);
        for (int k = 0; k < 8; k++) x[k] = h64::from_u128(k + 3);
        t0 = now();
        for (int i = 0; i < 4000000; i++) for (int k = 0; k < 8; k++) x[k] = h64::mul_portable(x[k], b);
        t1 = now();
        for (int k = 1; k < 8; k++) x[0] = h64::add(x[0], x[k]);
        h64::to_bytes(o, x[0]);
        printf("Fr mul_portable:            %.2f ns %02x\n", (t1 - t0) / 3.2e7, o[0]);
        t0 = now();
        for (int i = 0; i < 4000000; i++) for (int k = 0; k < 8; k++) x[k] = h64::add(x[k], b);
        t1 = now();
        h64::to_bytes(o, x[3]);
        printf("Fr add:                     %.2f ns %02x\n", (t1 - t0) / 3.2e7, o[0]);
    }
    std::vector<uint8_t> m(21000, 7), m2(21003, 9);
    uint8_t d[32], e[32], pre[5] = {1, 2, 3, 4, 5};
    double t0 = now();
    for (int i = 0; i < 2000; i++) { m[0] = i; sha::digest3(d, m.data(), m.size(), 0, 0, 0, 0); }
    double t1 = now();
    printf("SHA-256 one stream:  %.3f ns/byte %02x\n", (t1 - t0) / 2000 / 21000, d[0]);
    t0 = now();
    for (int i = 0; i < 2000; i++) { m[0] = i; sha::digest2x2(d, pre, 5, m.data(), m.size(), e, pre, 4, m2.data(), m2.size()); }
    t1 = now();
    printf("SHA-256 two streams: %.3f ns/byte %02x\n", (t1 - t0) / 2000 / 42000, d[0] ^ e[0]);
}
