// oracle/shim_stream_driver.cpp -- TEST INFRASTRUCTURE: a ten-line application of the libzpaq stream API
// (libzpaq::compress / libzpaq::decompress over Reader/Writer, libzpaq.h:866-890), compiled against the UNCHANGED
// libzpaq.h and linked like the zpaq CLI: shim (zpaq_b200/csrc/libzpaq_shim.cpp) + reference object + libzpq_b200.so.
//   shim_stream c METHOD in out     compress file `in` into archive `out`
//   shim_stream d - in out          decompress archive `in` into `out`
//   shim_stream D N in out          the same block by block through Decompresser (findBlock / findFilename / readComment /
//                                   decompress(N) in pieces of N bytes / readSegmentEnd), checking every stored SHA-1
#include "libzpaq.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void libzpaq::error(const char* msg) {      // the application supplies it; it must not return (libzpaq.h:857)
  fprintf(stderr, "shim_stream: %s\n", msg);
  exit(2);
}

struct FileReader : public libzpaq::Reader {
  FILE* f;
  int get() { return getc(f); }
  int read(char* buf, int n) { return (int)fread(buf, 1, n, f); }
};
struct FileWriter : public libzpaq::Writer {
  FILE* f;
  void put(int c) { putc(c, f); }
  void write(const char* buf, int n) { fwrite(buf, 1, n, f); }
};

int main(int argc, char** argv) {
  if (argc != 5) { fprintf(stderr, "usage: shim_stream c|d METHOD in out\n"); return 1; }
  FileReader in; FileWriter out;
  in.f = fopen(argv[3], "rb"); out.f = fopen(argv[4], "wb");
  if (!in.f || !out.f) { perror("open"); return 1; }
  if (argv[1][0] == 'c') libzpaq::compress(&in, &out, argv[2], "name", "comment", true);
  else if (argv[1][0] == 'D') {
    // what zpaq's extract threads do (zpaq.cpp:2848-2867), one block at a time
    const int piece = atoi(argv[2]) > 0 ? atoi(argv[2]) : 65536;
    libzpaq::Decompresser d;
    d.setInput(&in);
    int nseg = 0;
    while (d.findBlock()) {
      while (d.findFilename()) {
        d.readComment();
        libzpaq::SHA1 sha;
        d.setOutput(&out);
        d.setSHA1(&sha);
        while (d.decompress(piece)) { }
        char want[21];
        d.readSegmentEnd(want);
        if (want[0] == 1 && memcmp(want + 1, sha.result(), 20) != 0) libzpaq::error("checksum mismatch");
        ++nseg;
      }
    }
    fprintf(stderr, "shim_stream: %d segments\n", nseg);
  }
  else libzpaq::decompress(&in, &out);
  fclose(out.f);
  return 0;
}
