// exhaustive-ish check of str_find_w against str_find (host build of the device runtime)
#define SDQLB200_EMU
#include "sdqlb200_rt.cuh"
#include <cstdio>
#include <cstdlib>
// wcsstr semantics bounded to the row (varchar.h:84-97): first occurrence before the first NUL
static int naive_find(const unsigned char* s, int w, const char* pat, int plen) {
    int n = 0;
    while (n < w && s[n]) ++n;
    for (int i = 0; i + plen <= n; ++i)
        if (!memcmp(s + i, pat, plen)) return i;
    return -1;
}
int main() {
    unsigned char buf[16 + 1024 * 8 + 16];
    srand(7);
    long bad = 0, n = 0;
    const char* pats[] = {"special", "requests", "green", "ab", "a", "Customer", "Complaints"};
    for (int it = 0; it < 300000; ++it) {
        int w = 5 + rand() % 100;
        int off = rand() % 8;
        unsigned char* s = buf + 16 + off;
        int len = rand() % (w + 1);
        const char* alpha = "abspecilrqutgnCmo ";
        for (int i = 0; i < w; ++i) s[i] = i < len ? alpha[rand() % 18] : 0;
        for (int i = -16; i < 0; ++i) s[i] = alpha[rand() % 18];      // neighbouring rows
        for (int i = w; i < w + 8; ++i) s[i] = alpha[rand() % 18];
        const char* p = pats[rand() % 7];
        int plen = (int)strlen(p);
        if (rand() % 3 == 0 && len >= plen) memcpy(s + rand() % (len - plen + 1), p, plen);  // plant a match
        int r = naive_find(s, w, p, plen);
        int a = sdqlrt::str_find(s, w, p, plen), b = sdqlrt::str_find_bytes(s, w, p, plen);
        ++n;
        if (a != r) { a = -99; }
        if (a != b || a != r) { if (bad < 5) printf("MISMATCH w=%d len=%d pat=%s ref=%d got=%d\n", w, len, p, a, b); ++bad; }
    }
    printf("%ld cases, %ld mismatches\n", n, bad);
    return bad != 0;
}
