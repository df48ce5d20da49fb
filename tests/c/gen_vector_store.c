/* gen_vector_store.c — TEST TOOL: writes a llamaindex-style vector_store.json quickly (Python's json needs minutes for
 * 10^8 numbers):  gen <path> <rows> <dim> <extra>   — `rows` document nodes followed by `extra` memory nodes, the way
 * index.insert appends them; values are deterministic in (row, col), so a file with extra = 1 starts with exactly the
 * bytes of the file with extra = 0 up to the last document embedding. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  FILE* f = fopen(argv[1], "w");
  if (!f) return 1;
  const long rows = atol(argv[2]), dim = atol(argv[3]), extra = atol(argv[4]);
  static char buf[1 << 20];
  setvbuf(f, buf, _IOFBF, sizeof buf);
  fputs("{\"embeddingDict\":{", f);
  for (long r = 0; r < rows + extra; r++) {
    fprintf(f, "%s\"%s-%ld\":[", r ? "," : "", r < rows ? "node" : "memory", r);
    for (long c = 0; c < dim; c++) {
      const float v = (float)((double)(mix((uint64_t)r * 8191u + (uint64_t)c) >> 40) / 16777216.0 - 0.5);
      fprintf(f, c ? ",%.9g" : "%.9g", (double)v);
    }
    fputc(']', f);
  }
  fputs("},\"textIdToRefDocId\":{},\"metadataDict\":{", f);
  for (long r = rows; r < rows + extra; r++)
    fprintf(f, "%s\"memory-%ld\":{\"type\":\"memory\",\"memoryId\":\"m%ld\"}", r > rows ? "," : "", r, r);
  fputs("}}\n", f);
  return fclose(f) != 0;
}
