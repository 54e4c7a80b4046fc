// java_shim.h -- TEST INFRASTRUCTURE.  The slice of the Java runtime that the reference's octree builder and SDF brush
// (src/engine/Octree.java, OctreeThread.java, Util.java, sdf/*.java) touch, so that their TEXT -- rewritten
// mechanically by oracle/build_ref_java.py, never copied into the repository -- compiles as C++ (oracle/_ref/).
// Semantics follow the Java language / library specification where C++ differs:
//   * byte / short / long are int8_t / int16_t / int64_t; compound assignment narrows silently (gcc wraps);
//   * arrays are references with shared ownership (jarray<T>): copying an int[] copies the reference;
//   * objects are plain pointers made by `new` and never freed one by one (ByteBuffers live in an arena the harness
//     empties after every chunk);
//   * (int) of a floating value saturates and maps NaN to 0 (JLS 5.1.3): J::to_int;
//   * Math.round(double) = floor(x + 0.5) as long; Math.pow / sqrt / ceil are the correctly rounded libm calls for the
//     arguments used here (squares, exact halves);
//   * ByteBuffer: absolute get/put, big-endian getInt/putInt/getShort/putShort (Octree.java:66 sets BIG_ENDIAN),
//     position/limit and the relative bulk put(ByteBuffer) of the chunk splice (Octree.java:333-335);
//   * Thread.start() runs run() at once (the eight OctreeThreads write to eight separate buffers: order-free).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <vector>

#include <sys/mman.h>

namespace javaref {

typedef int8_t jbyte;
typedef int16_t jshort;
typedef int64_t jlong;

template <class T>
struct jarray {
  std::shared_ptr<std::vector<T>> v;
  jarray() {}
  jarray(std::nullptr_t) {}
  jarray(std::initializer_list<T> il) : v(std::make_shared<std::vector<T>>(il)) {}
  static jarray make(int n) {
    jarray a;
    a.v = std::make_shared<std::vector<T>>((size_t)n);
    return a;
  }
  T &operator[](int i) { return (*v).at((size_t)i); }
  const T &operator[](int i) const { return (*v).at((size_t)i); }
  typename std::vector<T>::iterator begin() { return v->begin(); }
  typename std::vector<T>::iterator end() { return v->end(); }
};
template <class T>
jarray<jarray<T>> make2(int a, int b) {
  jarray<jarray<T>> r = jarray<jarray<T>>::make(a);
  for (int i = 0; i < a; i++) r[i] = jarray<T>::make(b);
  return r;
}

struct J {
  template <class T>
  static int to_int(T x) {
    if constexpr (std::is_floating_point<T>::value) {
      if (x != x) return 0;
      if (x >= 2147483647.0) return 2147483647;
      if (x <= -2147483648.0) return (int)(-2147483647 - 1);
      return (int)x;
    } else {
      return (int)x;
    }
  }
};

struct Math {
  static int abs(int a) { return a < 0 ? -a : a; }
  static int max(int a, int b) { return a > b ? a : b; }
  static int min(int a, int b) { return a < b ? a : b; }
  static double pow(double a, double b) { return std::pow(a, b); }
  static double sqrt(double a) { return std::sqrt(a); }
  static double ceil(double a) { return std::ceil(a); }
  static jlong round(double a) {
    if (a != a) return 0;
    return (jlong)std::floor(a + 0.5);
  }
};

struct System {
  static double currentTimeMillis() { return 0.0; }
};

struct ByteOrder {
  enum Kind { BIG_ENDIAN_, LITTLE_ENDIAN_ };
};

struct ByteBuffer {
  uint8_t *data = nullptr;
  size_t capacity = 0, pos = 0, lim = 0;
  bool owned = false;
  jbyte get(int i) const { return (jbyte)data[check(i, 1)]; }
  ByteBuffer *put(int i, jbyte b) {
    data[check(i, 1)] = (uint8_t)b;
    return this;
  }
  int getInt(int i) const {
    const uint8_t *p = data + check(i, 4);
    return (int)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]);
  }
  ByteBuffer *putInt(int i, int v) {
    uint8_t *p = data + check(i, 4);
    p[0] = (uint8_t)((uint32_t)v >> 24);
    p[1] = (uint8_t)((uint32_t)v >> 16);
    p[2] = (uint8_t)((uint32_t)v >> 8);
    p[3] = (uint8_t)v;
    return this;
  }
  jshort getShort(int i) const {
    const uint8_t *p = data + check(i, 2);
    return (jshort)(uint16_t)(((uint32_t)p[0] << 8) | (uint32_t)p[1]);
  }
  ByteBuffer *putShort(int i, jshort v) {
    uint8_t *p = data + check(i, 2);
    p[0] = (uint8_t)((uint16_t)v >> 8);
    p[1] = (uint8_t)v;
    return this;
  }
  ByteBuffer *order(ByteOrder::Kind) { return this; }  // big-endian is the only order the builder asks for
  ByteBuffer *position(int p) {
    pos = (size_t)p;
    return this;
  }
  ByteBuffer *limit(int l) {
    lim = (size_t)l;
    return this;
  }
  ByteBuffer *put(ByteBuffer *src) {  // relative bulk put: src.remaining() bytes from src.position() to this.position()
    const size_t n = src->lim - src->pos;
    if (pos + n > capacity) throw std::out_of_range("BufferOverflowException");
    std::memcpy(data + pos, src->data + src->pos, n);
    pos += n;
    src->pos += n;
    return this;
  }
  size_t check(int i, int n) const {
    if (i < 0 || (size_t)i + (size_t)n > capacity) throw std::out_of_range("IndexOutOfBoundsException");
    return (size_t)i;
  }
};

// BufferUtils.createByteBuffer: zero-filled, off-heap.  mmap so that the reference's own sizes (2 GB for the world, 125 MB per
// OctreeThread, Constants.java) cost only the pages that are touched; the harness frees everything made since a mark.
struct BufferUtils {
  static std::vector<ByteBuffer *> &arena() {
    static std::vector<ByteBuffer *> a;
    return a;
  }
  static ByteBuffer *createByteBuffer(int bytes) {
    ByteBuffer *b = new ByteBuffer();
    b->capacity = (size_t)bytes;
    b->lim = b->capacity;
    void *p = mmap(nullptr, b->capacity ? b->capacity : 1, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) abort();
    b->data = (uint8_t *)p;
    b->owned = true;
    arena().push_back(b);
    return b;
  }
  static size_t mark() { return arena().size(); }
  static void release_to(size_t mark) {
    while (arena().size() > mark) {
      ByteBuffer *b = arena().back();
      arena().pop_back();
      munmap(b->data, b->capacity ? b->capacity : 1);
      delete b;
    }
  }
};

template <class T>
struct ArrayList {
  std::vector<T> v;
  void add(T x) { v.push_back(x); }
  int size() const { return (int)v.size(); }
  typename std::vector<T>::iterator begin() { return v.begin(); }
  typename std::vector<T>::iterator end() { return v.end(); }
};

template <class T>
using Consumer = std::function<void(T)>;

struct Thread {
  virtual ~Thread() {}
  virtual void run() {}
  void start() { run(); }
  bool isAlive() { return false; }
};

}  // namespace javaref
