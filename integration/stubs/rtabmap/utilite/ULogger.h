#pragma once
#include <cstdio>
#define UERROR(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define UWARN(...) UERROR(__VA_ARGS__)
