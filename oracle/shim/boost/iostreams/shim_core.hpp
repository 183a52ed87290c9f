// Minimal stand-in for the few boost::iostreams pieces the SPRING sources use
// (gzip filter + file/back_inserter sinks + filtering_{ostream,streambuf}).
// TEST INFRASTRUCTURE ONLY: lets oracle/Makefile compile the reference's own
// sources in place (under /root/reference) with plain g++ + zlib, without
// running the reference's cmake/boost build.  Not part of the product.
#ifndef SPRING_B200_ORACLE_BOOST_SHIM_CORE_HPP
#define SPRING_B200_ORACLE_BOOST_SHIM_CORE_HPP
#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <streambuf>
#include <string>
#include <vector>

namespace boost {
namespace iostreams {

struct input {};
struct output {};

struct gzip_params {
  int level;
  gzip_params(int l = Z_DEFAULT_COMPRESSION) : level(l) {}
};
struct gzip_compressor {
  int level;
  gzip_compressor(const gzip_params &p = gzip_params()) : level(p.level) {}
};
struct gzip_decompressor {};

struct file_sink {
  std::string path;
  std::ios_base::openmode mode;
  file_sink(const std::string &p,
            std::ios_base::openmode m = std::ios_base::out)
      : path(p), mode(m) {}
};

struct string_back_inserter {
  std::string *target;
};
inline string_back_inserter back_inserter(std::string &s) {
  string_back_inserter b;
  b.target = &s;
  return b;
}

namespace shim_detail {

// streambuf: bytes in -> (optional gzip deflate) -> file or std::string
class out_buf : public std::streambuf {
 public:
  out_buf() : gz_(false), level_(Z_DEFAULT_COMPRESSION), file_(NULL),
              str_(NULL), open_(false), zinit_(false) {
    setp(ibuf_, ibuf_ + sizeof(ibuf_));
  }
  ~out_buf() { finish(); }
  void set_gzip(int level) { gz_ = true; level_ = level; }
  void set_file(const std::string &p) {
    file_ = std::fopen(p.c_str(), "wb");
    if (!file_) throw std::runtime_error("boost shim: cannot open " + p);
    start();
  }
  void set_string(std::string *s) { str_ = s; start(); }
  void finish() {
    if (!open_) return;
    flush_in(true);
    if (zinit_) { deflateEnd(&zs_); zinit_ = false; }
    if (file_) { std::fclose(file_); file_ = NULL; }
    str_ = NULL;
    open_ = false;
  }

 protected:
  int_type overflow(int_type ch) {
    flush_in(false);
    if (ch != traits_type::eof()) { *pptr() = (char)ch; pbump(1); }
    return traits_type::not_eof(ch);
  }
  int sync() { flush_in(false); return 0; }

 private:
  void start() {
    open_ = true;
    if (gz_) {
      std::memset(&zs_, 0, sizeof(zs_));
      if (deflateInit2(&zs_, level_, Z_DEFLATED, 15 + 16, 8,
                       Z_DEFAULT_STRATEGY) != Z_OK)
        throw std::runtime_error("boost shim: deflateInit2 failed");
      zinit_ = true;
    }
  }
  void emit(const char *p, size_t n) {
    if (!n) return;
    if (file_) std::fwrite(p, 1, n, file_);
    else if (str_) str_->append(p, n);
  }
  void flush_in(bool last) {
    size_t n = pptr() - pbase();
    if (!open_) { setp(ibuf_, ibuf_ + sizeof(ibuf_)); return; }
    if (!gz_) {
      emit(pbase(), n);
    } else {
      zs_.next_in = (Bytef *)pbase();
      zs_.avail_in = (uInt)n;
      int flush = last ? Z_FINISH : Z_NO_FLUSH;
      do {
        zs_.next_out = (Bytef *)obuf_;
        zs_.avail_out = sizeof(obuf_);
        deflate(&zs_, flush);
        emit(obuf_, sizeof(obuf_) - zs_.avail_out);
      } while (zs_.avail_out == 0 || zs_.avail_in > 0);
    }
    setp(ibuf_, ibuf_ + sizeof(ibuf_));
  }
  bool gz_;
  int level_;
  std::FILE *file_;
  std::string *str_;
  bool open_, zinit_;
  z_stream zs_;
  char ibuf_[1 << 16];
  char obuf_[1 << 16];
};

// streambuf: std::istream (compressed) -> (optional gzip inflate) -> bytes out
class in_buf : public std::streambuf {
 public:
  in_buf() : gz_(false), src_(NULL), zinit_(false), eof_(false) {
    setg(obuf_, obuf_, obuf_);
  }
  ~in_buf() { if (zinit_) inflateEnd(&zs_); }
  void set_gzip() { gz_ = true; }
  void set_source(std::istream *s) {
    src_ = s;
    if (gz_) {
      std::memset(&zs_, 0, sizeof(zs_));
      if (inflateInit2(&zs_, 15 + 32) != Z_OK)
        throw std::runtime_error("boost shim: inflateInit2 failed");
      zinit_ = true;
    }
  }

 protected:
  int_type underflow() {
    if (gptr() < egptr()) return traits_type::to_int_type(*gptr());
    if (!src_ || eof_) return traits_type::eof();
    size_t produced = 0;
    if (!gz_) {
      src_->read(obuf_, sizeof(obuf_));
      produced = (size_t)src_->gcount();
      if (produced == 0) eof_ = true;
    } else {
      while (produced == 0 && !eof_) {
        if (zs_.avail_in == 0) {
          src_->read(ibuf_, sizeof(ibuf_));
          zs_.next_in = (Bytef *)ibuf_;
          zs_.avail_in = (uInt)src_->gcount();
        }
        zs_.next_out = (Bytef *)obuf_;
        zs_.avail_out = sizeof(obuf_);
        int rc = inflate(&zs_, Z_NO_FLUSH);
        produced = sizeof(obuf_) - zs_.avail_out;
        if (rc == Z_STREAM_END) {
          // concatenated gzip members: restart if more input follows
          if (zs_.avail_in == 0) {
            src_->read(ibuf_, sizeof(ibuf_));
            zs_.next_in = (Bytef *)ibuf_;
            zs_.avail_in = (uInt)src_->gcount();
          }
          if (zs_.avail_in == 0) eof_ = true;
          else inflateReset(&zs_);
        } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
          eof_ = true;
        } else if (rc == Z_BUF_ERROR && zs_.avail_in == 0 && src_->eof()) {
          eof_ = true;
        }
      }
    }
    if (produced == 0) return traits_type::eof();
    setg(obuf_, obuf_, obuf_ + produced);
    return traits_type::to_int_type(*gptr());
  }

 private:
  bool gz_;
  std::istream *src_;
  bool zinit_, eof_;
  z_stream zs_;
  char ibuf_[1 << 16];
  char obuf_[1 << 16];
};

}  // namespace shim_detail

class filtering_ostream : public std::ostream {
 public:
  filtering_ostream() : std::ostream(NULL) { rdbuf(&buf_); }
  ~filtering_ostream() { buf_.finish(); }
  void push(const gzip_compressor &c) { buf_.set_gzip(c.level); }
  void push(const file_sink &f) { buf_.set_file(f.path); }
  void push(const string_back_inserter &b) { buf_.set_string(b.target); }
  void pop() { flush(); buf_.finish(); }
  void reset() { flush(); buf_.finish(); }

 private:
  shim_detail::out_buf buf_;
};

inline void close(filtering_ostream &o) { o.reset(); }

template <typename Mode>
class filtering_streambuf : public shim_detail::in_buf {
 public:
  void push(const gzip_decompressor &) { set_gzip(); }
  void push(std::istream &s) { set_source(&s); }
};

}  // namespace iostreams
}  // namespace boost
#endif
