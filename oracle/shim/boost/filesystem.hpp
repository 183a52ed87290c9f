// Minimal stand-in for the boost::filesystem calls the SPRING sources make
// (path, directory_iterator, file_size, exists, create_directory, remove_all).
// TEST INFRASTRUCTURE ONLY (see iostreams/shim_core.hpp).
#ifndef SPRING_B200_ORACLE_BOOST_SHIM_FILESYSTEM_HPP
#define SPRING_B200_ORACLE_BOOST_SHIM_FILESYSTEM_HPP
#include <dirent.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace filesystem {

class path {
 public:
  path() {}
  path(const std::string &s) : s_(s) {}
  path(const char *s) : s_(s) {}
  const std::string &string() const { return s_; }
  path filename() const {
    size_t p = s_.find_last_of('/');
    return path(p == std::string::npos ? s_ : s_.substr(p + 1));
  }
 private:
  std::string s_;
};

class directory_entry {
 public:
  directory_entry() {}
  explicit directory_entry(const std::string &p) : p_(p) {}
  const filesystem::path &path() const { return p_; }
 private:
  filesystem::path p_;
};

class directory_iterator {
 public:
  directory_iterator() : idx_(0) {}
  explicit directory_iterator(const path &p) : idx_(0) {
    entries_ = std::make_shared<std::vector<directory_entry> >();
    DIR *d = opendir(p.string().c_str());
    if (!d) throw std::runtime_error("boost shim: cannot open dir " + p.string());
    while (struct dirent *e = readdir(d)) {
      std::string n = e->d_name;
      if (n == "." || n == "..") continue;
      entries_->push_back(directory_entry(p.string() + "/" + n));
    }
    closedir(d);
    if (entries_->empty()) entries_.reset();
  }
  directory_iterator &operator++() {
    if (entries_ && ++idx_ >= entries_->size()) { entries_.reset(); idx_ = 0; }
    return *this;
  }
  const directory_entry &operator*() const { return (*entries_)[idx_]; }
  const directory_entry *operator->() const { return &(*entries_)[idx_]; }
  bool operator==(const directory_iterator &o) const {
    return entries_ == o.entries_ && idx_ == o.idx_;
  }
  bool operator!=(const directory_iterator &o) const { return !(*this == o); }
 private:
  std::shared_ptr<std::vector<directory_entry> > entries_;
  size_t idx_;
};

inline uintmax_t file_size(const path &p) {
  struct stat st;
  if (stat(p.string().c_str(), &st) != 0)
    throw std::runtime_error("boost shim: stat failed " + p.string());
  return (uintmax_t)st.st_size;
}
inline bool exists(const path &p) {
  struct stat st;
  return stat(p.string().c_str(), &st) == 0;
}
inline bool create_directory(const path &p) {
  return mkdir(p.string().c_str(), 0777) == 0;
}
inline uintmax_t remove_all(const path &p) {
  struct stat st;
  if (lstat(p.string().c_str(), &st) != 0) return 0;
  uintmax_t n = 0;
  if (S_ISDIR(st.st_mode)) {
    DIR *d = opendir(p.string().c_str());
    if (d) {
      while (struct dirent *e = readdir(d)) {
        std::string nm = e->d_name;
        if (nm == "." || nm == "..") continue;
        n += remove_all(path(p.string() + "/" + nm));
      }
      closedir(d);
    }
    rmdir(p.string().c_str());
  } else {
    unlink(p.string().c_str());
  }
  return n + 1;
}

}  // namespace filesystem
}  // namespace boost
#endif
