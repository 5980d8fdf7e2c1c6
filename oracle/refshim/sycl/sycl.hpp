/*
 * sycl/sycl.hpp — TEST-ONLY host shim of the small part of SYCL 2020 the reference's hot-path sources
 * use, so that felipeagc/sycl-ray-tracer's OWN unmodified files (src/xorshift.hpp, camera.hpp,
 * util.hpp, material.hpp, trace_ray.hpp, render_megakernel.cpp, render_wavefront.cpp) compile with
 * plain g++ and RUN on the CPU to pin the oracle (oracle/Makefile `ref`, output oracle/_ref/).
 * This is our code, not a copy of any SYCL implementation: kernels are executed serially, one
 * work-group at a time, every work-item on its own ucontext fiber so that group barriers behave.
 * Numeric definitions the SYCL/OpenCL runtime would supply are the ones DESIGN.md pins
 * (normalize = v * (1/sqrt(dot)), nearest/repeat texel rule, unorm8 = sat(rte(f*255))).
 */
#pragma once
#include <ucontext.h>

#include <array>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <exception>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <optional>
#include <string>
#include <type_traits>
#include <vector>

namespace sycl {

/* ------------------------------------------------------------------ half / vec */
struct half {
    _Float16 v;
    half() = default;
    half(float f) : v((_Float16)f) {}
    operator float() const { return (float)v; }
};

template <class T, int N>
struct alignas(N == 3 ? 4 * sizeof(T) : N * sizeof(T)) vec {
    T s[N == 3 ? 4 : N];
    vec() = default;
    constexpr vec(T a) : s{} {
        for (int i = 0; i < N; i++) s[i] = a;
    }
    template <int M = N, std::enable_if_t<M == 2, int> = 0>
    constexpr vec(T a, T b) : s{a, b} {}
    template <int M = N, std::enable_if_t<M == 3, int> = 0>
    constexpr vec(T a, T b, T c) : s{a, b, c, T()} {}
    template <int M = N, std::enable_if_t<M == 4, int> = 0>
    constexpr vec(T a, T b, T c, T d) : s{a, b, c, d} {}
    template <int M = N, std::enable_if_t<M == 4, int> = 0>
    constexpr vec(const vec<T, 3> &a, T d) : s{a.s[0], a.s[1], a.s[2], d} {}
    constexpr T &operator[](int i) { return s[i]; }
    constexpr const T &operator[](int i) const { return s[i]; }
    constexpr T x() const { return s[0]; }
    constexpr T y() const { return s[1]; }
    constexpr T z() const { return s[2]; }
    constexpr T w() const { return s[3]; }
    template <class U>
    vec<U, N> convert() const {
        vec<U, N> r;
        for (int i = 0; i < N; i++) r.s[i] = U((float)s[i]);
        return r;
    }
    vec &operator+=(const vec &o) {
        for (int i = 0; i < N; i++) s[i] = s[i] + o.s[i];
        return *this;
    }
    vec &operator*=(T o) {
        for (int i = 0; i < N; i++) s[i] = s[i] * o;
        return *this;
    }
    vec &operator/=(T o) {
        for (int i = 0; i < N; i++) s[i] = s[i] / o;
        return *this;
    }
};
#define RS_BIN(op)                                                        \
    template <class T, int N>                                             \
    vec<T, N> operator op(const vec<T, N> &a, const vec<T, N> &b) {       \
        vec<T, N> r;                                                      \
        for (int i = 0; i < N; i++) r.s[i] = a.s[i] op b.s[i];            \
        return r;                                                         \
    }                                                                     \
    template <class T, int N, class S, std::enable_if_t<std::is_arithmetic_v<S>, int> = 0> \
    vec<T, N> operator op(const vec<T, N> &a, S b) {                      \
        vec<T, N> r;                                                      \
        for (int i = 0; i < N; i++) r.s[i] = a.s[i] op(T) b;              \
        return r;                                                         \
    }                                                                     \
    template <class T, int N, class S, std::enable_if_t<std::is_arithmetic_v<S>, int> = 0> \
    vec<T, N> operator op(S a, const vec<T, N> &b) {                      \
        vec<T, N> r;                                                      \
        for (int i = 0; i < N; i++) r.s[i] = (T)a op b.s[i];              \
        return r;                                                         \
    }
RS_BIN(+)
RS_BIN(-)
RS_BIN(*)
RS_BIN(/)
#undef RS_BIN
template <class T, int N>
vec<T, N> operator-(const vec<T, N> &a) {
    vec<T, N> r;
    for (int i = 0; i < N; i++) r.s[i] = -a.s[i];
    return r;
}
using float2 = vec<float, 2>;
using float3 = vec<float, 3>;
using float4 = vec<float, 4>;
using int2 = vec<int, 2>;
using half3 = vec<half, 3>;

/* math: association orders are the ones pinned in DESIGN.md */
inline float sqrt(float a) { return std::sqrt(a); }
inline float fabs(float a) { return std::fabs(a); }
inline float pow(float a, float b) {
    if (b == 5.0f) { /* pinned expansion of pow(x, 5) */
        float x2 = a * a;
        return (x2 * x2) * a;
    }
    return std::pow(a, b);
}
inline float dot(const float3 &a, const float3 &b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline float length(const float3 &a) { return std::sqrt(dot(a, a)); }
inline float3 cross(const float3 &a, const float3 &b) {
    return float3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline float3 normalize(const float3 &a) {
    float inv = 1.0f / std::sqrt(dot(a, a));
    return a * inv;
}
inline float3 clamp(const float3 &a, float lo, float hi) {
    return float3(std::fmin(std::fmax(a[0], lo), hi), std::fmin(std::fmax(a[1], lo), hi), std::fmin(std::fmax(a[2], lo), hi));
}

/* ------------------------------------------------------------------ ranges / ids */
template <int N>
struct range {
    size_t v[N];
    range() : v{} {}
    range(size_t a) : v{} {
        static_assert(N >= 1);
        for (int i = 0; i < N; i++) v[i] = a;
    }
    template <int M = N, std::enable_if_t<M == 2, int> = 0>
    range(size_t a, size_t b) : v{a, b} {}
    template <int M = N, std::enable_if_t<M == 3, int> = 0>
    range(size_t a, size_t b, size_t c) : v{a, b, c} {}
    size_t &operator[](int i) { return v[i]; }
    size_t operator[](int i) const { return v[i]; }
    size_t size() const {
        size_t s = 1;
        for (int i = 0; i < N; i++) s *= v[i];
        return s;
    }
};
template <>
struct range<1> {
    size_t v[1];
    range() : v{} {}
    range(size_t a) : v{a} {}
    size_t &operator[](int i) { return v[i]; }
    size_t operator[](int i) const { return v[i]; }
    size_t size() const { return v[0]; }
    operator size_t() const { return v[0]; }
};
template <int N>
range<N> operator*(const range<N> &a, const range<N> &b) {
    range<N> r;
    for (int i = 0; i < N; i++) r.v[i] = a.v[i] * b.v[i];
    return r;
}
template <int N>
range<N> operator+(const range<N> &a, const range<N> &b) {
    range<N> r;
    for (int i = 0; i < N; i++) r.v[i] = a.v[i] + b.v[i];
    return r;
}
template <int N>
struct id {
    size_t v[N];
    id() : v{} {}
    size_t operator[](int i) const { return v[i]; }
};
template <>
struct id<1> {
    size_t v[1];
    id() : v{} {}
    size_t operator[](int i) const { return v[i]; }
    operator size_t() const { return v[0]; }
};
template <int N>
struct item {
    id<N> i;
    size_t operator[](int k) const { return i.v[k]; }
};
template <int N>
struct nd_range {
    range<N> global, local;
    nd_range(range<N> g, range<N> l) : global(g), local(l) {}
};

namespace access {
enum class mode { read, write, read_write };
enum class target { device, image, image_array, local };
enum class fence_space { local_space, global_space, global_and_local };
enum class address_space { global_space, local_space };
} // namespace access
using access_mode = access::mode;
enum memory_order { memory_order_relaxed };
enum memory_scope { memory_scope_device, memory_scope_work_group };

/* fibers: one per work-item so that nd_item::barrier() really waits for the whole group */
namespace detail {
struct FiberGroup {
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<std::unique_ptr<char[]>> stacks;
    std::vector<char> done;
    std::function<void(size_t)> body;
    size_t current = 0;
    static FiberGroup *&active() {
        static FiberGroup *g = nullptr;
        return g;
    }
    static void entry() {
        FiberGroup *g = active();
        const size_t me = g->current;
        g->body(me);
        g->done[me] = 1;
        swapcontext(&g->ctx[me], &g->sched);
    }
    void yield() {
        const size_t me = current;
        swapcontext(&ctx[me], &sched);
    }
    void run(size_t n, std::function<void(size_t)> f) {
        const size_t kStack = 256 * 1024;
        body = std::move(f);
        if (ctx.size() < n) {
            ctx.resize(n);
            while (stacks.size() < n) stacks.emplace_back(new char[kStack]);
        }
        done.assign(n, 0);
        active() = this;
        for (size_t i = 0; i < n; i++) {
            getcontext(&ctx[i]);
            ctx[i].uc_stack.ss_sp = stacks[i].get();
            ctx[i].uc_stack.ss_size = kStack;
            ctx[i].uc_link = &sched;
            makecontext(&ctx[i], (void (*)())entry, 0);
        }
        bool any = true;
        while (any) { /* round robin: every live item runs up to its next barrier (or its end) */
            any = false;
            for (size_t i = 0; i < n; i++) {
                if (done[i]) continue;
                current = i;
                swapcontext(&sched, &ctx[i]);
                if (!done[i]) any = true;
            }
        }
        active() = nullptr;
    }
};
} // namespace detail

struct kernel_handler {};
template <int N>
struct nd_item {
    size_t gid[N], lid[N], grange[N], lrange[N];
    id<N> get_global_id() const {
        id<N> r;
        for (int i = 0; i < N; i++) r.v[i] = gid[i];
        return r;
    }
    id<N> get_local_id() const {
        id<N> r;
        for (int i = 0; i < N; i++) r.v[i] = lid[i];
        return r;
    }
    size_t get_global_linear_id() const { /* dimension 0 is the slowest (SYCL 2020 4.9.1.5) */
        size_t r = 0;
        for (int i = 0; i < N; i++) r = r * grange[i] + gid[i];
        return r;
    }
    size_t get_local_linear_id() const {
        size_t r = 0;
        for (int i = 0; i < N; i++) r = r * lrange[i] + lid[i];
        return r;
    }
    void barrier(access::fence_space = access::fence_space::global_and_local) const {
        detail::FiberGroup::active()->yield();
    }
};

template <class T, memory_order O, memory_scope S, access::address_space A = access::address_space::global_space>
struct atomic_ref {
    T &r;
    explicit atomic_ref(T &x) : r(x) {}
    T operator=(T v) const {
        r = v;
        return v;
    }
    operator T() const { return r; }
    T fetch_add(T v) const {
        T o = r;
        r += v;
        return o;
    }
    T operator+=(T v) const { return r += v; }
};

/* ------------------------------------------------------------------ devices / queues */
namespace info {
namespace device {
struct name {
    using type = std::string;
};
} // namespace device
} // namespace info
struct device {
    device() = default;
    template <class Sel>
    explicit device(Sel) {}
    template <class I>
    typename I::type get_info() const { return "refshim host (serial work-groups, fibers)"; }
};
struct context {
    context() = default;
    explicit context(const device &) {}
};
struct exception : std::exception {
    const char *what() const noexcept override { return "sycl::exception (shim)"; }
};
using exception_list = std::vector<std::exception_ptr>;
struct event {
    void wait() {}
    void wait_and_throw() {}
};
struct handler;
struct queue {
    queue() = default;
    template <class H>
    queue(const device &, H) {}
    device get_device() const { return device(); }
    template <class F>
    event submit(F f);
    void wait() {}
    void wait_and_throw() {}
};
namespace usm {
enum class alloc { host, device, shared };
}
namespace ext::oneapi::property::usm {
struct device_read_only {};
} // namespace ext::oneapi::property::usm
template <class T>
T *malloc_shared(size_t n, const queue &) { return (T *)std::calloc(n ? n : 1, sizeof(T)); }
inline void free(void *p, const queue &) { std::free(p); }
inline void *aligned_alloc(size_t a, size_t bytes, const queue &, usm::alloc) {
    /* zero-filled: the reference never initialises the second queue's ray_buffer_length
     * (src/render_wavefront.hpp:19-21) and relies on fresh USM pages reading as zero */
    const size_t n = (bytes + 63) / 64 * 64;
    void *p = std::aligned_alloc(a < 16 ? 16 : a, n);
    if (p) std::memset(p, 0, n);
    return p;
}
template <class P>
void *aligned_alloc_shared(size_t a, size_t bytes, const queue &q, P) { return aligned_alloc(a, bytes, q, usm::alloc::shared); }
inline void *aligned_alloc_shared(size_t a, size_t bytes, const queue &q) { return aligned_alloc(a, bytes, q, usm::alloc::shared); }
inline void *aligned_alloc_device(size_t a, size_t bytes, const queue &q) { return aligned_alloc(a, bytes, q, usm::alloc::device); }

/* ------------------------------------------------------------------ images / samplers */
enum class image_channel_order { rgba };
enum class image_channel_type { unorm_int8, fp32 };
enum class coordinate_normalization_mode { normalized, unnormalized };
enum class addressing_mode { repeat, clamp };
enum class filtering_mode { nearest, linear };
struct sampler {
    sampler() = default;
    sampler(coordinate_normalization_mode, addressing_mode, filtering_mode) {}
};
template <int N>
struct image {
    void *data = nullptr;
    std::shared_ptr<std::vector<uint8_t>> own;
    image_channel_type type;
    range<N> size;
    image(void *host, image_channel_order, image_channel_type t, range<N> r) : data(host), type(t), size(r) {}
    image(image_channel_order, image_channel_type t, range<N> r) : type(t), size(r) {
        own = std::make_shared<std::vector<uint8_t>>(r.size() * (t == image_channel_type::fp32 ? 16 : 4), 0);
        data = own->data();
    }
    template <class T, access::mode M>
    auto get_access(handler &);
};
inline float4 rs_load_texel(const void *data, image_channel_type t, size_t idx) {
    if (t == image_channel_type::fp32) {
        const float *p = (const float *)data + idx * 4;
        return float4(p[0], p[1], p[2], p[3]);
    }
    const uint8_t *p = (const uint8_t *)data + idx * 4;
    return float4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f);
}
inline uint8_t rs_unorm8(float f) { /* OpenCL 3.0 8.3.1.1: convert_uchar_sat_rte(f * 255.0f) */
    float c = f * 255.0f;
    if (!(c > 0.0f)) return 0;
    if (c >= 255.0f) return 255;
    return (uint8_t)std::nearbyint(c);
}
inline int rs_wrap(float s, int w) { /* nearest + repeat + normalised (OpenCL 3.0 8.2) */
    float u = (s - std::floor(s)) * (float)w;
    int i = (int)std::floor(u);
    if (i > w - 1) i -= w;
    if (i < 0) i = 0;
    return i;
}
template <class T, int N, access::mode M, access::target Tg = access::target::image>
struct accessor;
template <access::mode M>
struct accessor<float4, 2, M, access::target::image> {
    image<2> *img;
    float4 read(const int2 &c) const { return rs_load_texel(img->data, img->type, (size_t)c[1] * img->size[0] + (size_t)c[0]); }
    void write(const int2 &c, const float4 &v) const {
        const size_t idx = (size_t)c[1] * img->size[0] + (size_t)c[0];
        if (img->type == image_channel_type::fp32) {
            float *p = (float *)img->data + idx * 4;
            for (int k = 0; k < 4; k++) p[k] = v[k];
        } else {
            uint8_t *p = (uint8_t *)img->data + idx * 4;
            for (int k = 0; k < 4; k++) p[k] = rs_unorm8(v[k]);
        }
    }
};
struct rs_image_slice {
    const image<3> *img;
    size_t layer;
    float4 read(const float2 &uv, const sampler &) const {
        const int w = (int)img->size[0], h = (int)img->size[1];
        const int ix = rs_wrap(uv[0], w), iy = rs_wrap(uv[1], h);
        return rs_load_texel(img->data, img->type, (layer * h + (size_t)iy) * w + (size_t)ix);
    }
};
template <>
struct accessor<float4, 2, access::mode::read, access::target::image_array> {
    const image<3> *img = nullptr;
    accessor() = default;
    accessor(image<3> &i, handler &) : img(&i) {}
    rs_image_slice operator[](size_t layer) const { return rs_image_slice{img, layer}; }
};
template <int N>
template <class T, access::mode M>
auto image<N>::get_access(handler &) {
    static_assert(N == 2);
    return accessor<float4, 2, M, access::target::image>{this};
}

/* ------------------------------------------------------------------ buffers / local memory */
template <class T>
struct rs_buffer_access {
    T *p;
    T &operator[](size_t i) const { return p[i]; }
};
template <class T>
struct buffer {
    T *p;
    buffer(T *host, size_t) : p(host) {}
    template <access::mode M>
    rs_buffer_access<T> get_access(handler &) { return {p}; }
    rs_buffer_access<T> get_host_access() { return {p}; }
};
template <class T, int N>
struct local_accessor {
    std::shared_ptr<std::vector<T>> mem;
    local_accessor(range<N> r, handler &) : mem(std::make_shared<std::vector<T>>(r.size())) {}
    T &operator[](size_t i) const {
#ifdef REFSHIM_CHECK
        if (i >= mem->size()) { fprintf(stderr, "local_accessor index %zu >= %zu\n", i, mem->size()); abort(); }
#endif
        return (*mem)[i];
    }
};
struct stream {
    stream(size_t, size_t, handler &) {}
};

/* ------------------------------------------------------------------ handler: serial execution */
struct handler {
    template <class K>
    void parallel_for(range<2> r, K k) {
        for (size_t a = 0; a < r[0]; a++)
            for (size_t b = 0; b < r[1]; b++) {
                if constexpr (std::is_invocable_v<K, item<2>>) {
                    item<2> it;
                    it.i.v[0] = a;
                    it.i.v[1] = b;
                    k(it);
                } else {
                    id<2> i;
                    i.v[0] = a;
                    i.v[1] = b;
                    k(i);
                }
            }
    }
    template <int N, class K>
    void parallel_for(nd_range<N> r, K k) {
        detail::FiberGroup fg;
        size_t groups[N], ng = 1, nl = 1;
        for (int i = 0; i < N; i++) {
            groups[i] = r.global[i] / r.local[i];
            ng *= groups[i];
            nl *= r.local[i];
        }
        for (size_t g = 0; g < ng; g++) {
            size_t gi[N], rem = g;
            for (int i = N - 1; i >= 0; i--) {
                gi[i] = rem % groups[i];
                rem /= groups[i];
            }
            fg.run(nl, [&](size_t l) {
                nd_item<N> it;
                size_t rl = l;
                for (int i = N - 1; i >= 0; i--) {
                    it.lid[i] = rl % r.local[i];
                    rl /= r.local[i];
                    it.gid[i] = gi[i] * r.local[i] + it.lid[i];
                    it.grange[i] = r.global[i];
                    it.lrange[i] = r.local[i];
                }
                if constexpr (std::is_invocable_v<K, nd_item<N>, kernel_handler>) k(it, kernel_handler{});
                else k(it);
            });
        }
    }
};
template <class F>
event queue::submit(F f) {
    handler h;
    f(h);
    return event{};
}

} // namespace sycl

/* Evaluation order of call arguments. The reference draws its random vectors as
 * `sycl::float3(rng(), rng(), rng())` (src/xorshift.hpp:26-36); the order of the three draws is
 * unspecified in C++. icpx (clang), the compiler the reference is built with, evaluates call
 * arguments left to right (x, y, z); g++ evaluates them right to left. To reproduce the
 * reference's actual behaviour with g++, every `float3( ... )` construction is rewritten into a
 * braced initialisation, whose left-to-right order the language guarantees. */
namespace sycl {
using float3_lr = vec<float, 3>;
}
using sycl::float3_lr;
#define float3(...) float3_lr{__VA_ARGS__}
