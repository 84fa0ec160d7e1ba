/* TEST INFRASTRUCTURE — plain-C restatement of the two hand-written image routines of the
 * reference's edge/label preparation, used by oracle/prep_cv2.py (which supplies the OpenCV
 * parts through the real cv2):
 *   Roberts  — /root/reference/csrc/DPE-MVS/DPE.cpp:9-25
 *   Connect  — /root/reference/csrc/DPE-MVS/DPE.cpp:28-134 (two-pass labelling of the zero
 *              set, 4-connectivity, with the reference's own parent-overwrite union)
 * Build: gcc -O2 -shared -fPIC oracle/prep_oracle.c -o oracle/_ref/libprep_oracle.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

void oracle_roberts(const unsigned char* src, int rows, int cols, unsigned char* dst) {
  for (int i = 0; i < rows; i++) {
    for (int j = 0; j < cols; j++) {
      int t1, t2;
      if (i > 0 && i < rows - 1 && j > 0 && j < cols - 1) {
        t1 = (int)src[i * cols + j] - (int)src[(i + 1) * cols + j + 1];
        t2 = (int)src[(i + 1) * cols + j] - (int)src[i * cols + j + 1];
      } else {
        t1 = t2 = 50;
      }
      /* (uchar)sqrt(...): out-of-range values wrap modulo 256 on x86 */
      dst[i * cols + j] = (unsigned char)((int)sqrt((double)(t1 * t1 + t2 * t2)) & 0xFF);
    }
  }
}

/* img: 0 = region pixel, 255 = edge pixel.  label: out, int32 rows*cols.  Returns the number
 * of labels (including label 0) and writes per-label pixel counts to cnt (caller allocates
 * rows*cols/2+2 ints). */
int oracle_connect(const unsigned char* img, int rows, int cols, int* label, int* cnt) {
  int cap = rows * cols / 2 + 2, n = 1;
  int* connection = (int*)malloc(sizeof(int) * (size_t)(rows * cols + 2));
  (void)cap;
  connection[0] = 0;
  for (int y = 0; y < rows; y++) {
    for (int x = 0; x < cols; x++) {
      const int c = y * cols + x;
      if (img[c] == 255) { label[c] = 0; continue; }
      const int left = (x > 0 && img[c] == 0 && img[c - 1] == 0);
      const int up = (y > 0 && img[c] == 0 && img[c - cols] == 0);
      if (left) label[c] = label[c - 1];
      if (up) label[c] = label[c - cols];
      if (!left && !up) {
        label[c] = n; connection[n] = n; n++;
      } else if (left && up) {
        const int ll = label[c - 1], ul = label[c - cols];
        if (ll > ul) { connection[ll] = ul; label[c] = ul; }
        else if (ll < ul) { connection[ul] = ll; label[c] = ll; }
      }
    }
  }
  for (int i = 1; i < n; i++) {
    int cur = connection[i], pre = connection[cur];
    while (pre != cur) { cur = pre; pre = connection[pre]; }
    connection[i] = cur;
  }
  int label_num = 1;
  int* mapping = (int*)calloc((size_t)n + 1, sizeof(int));
  for (int i = 1; i < n; i++)
    if (connection[i] == i) mapping[i] = label_num++;
  for (int i = 1; i < n; i++) connection[i] = mapping[connection[i]];
  memset(cnt, 0, sizeof(int) * (size_t)label_num);
  for (int i = 0; i < rows * cols; i++) { label[i] = connection[label[i]]; cnt[label[i]]++; }
  free(mapping); free(connection);
  return label_num;
}
