#pragma once
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <string>
namespace boost { namespace iostreams {
struct mapped_file_source {
  mapped_file_source() {}
  explicit mapped_file_source(std::string const& f) { open(f); }
  explicit mapped_file_source(const char* f) { open(f); }
  void open(std::string const& f) { int fd = ::open(f.c_str(), O_RDONLY); if (fd < 0) return; struct stat st; fstat(fd, &st); m_size = st.st_size; m_data = (const char*)mmap(0, m_size, PROT_READ, MAP_PRIVATE, fd, 0); ::close(fd); if (m_data == MAP_FAILED) m_data = nullptr; }
  bool is_open() const { return m_data != nullptr; }
  const char* data() const { return m_data; }
  size_t size() const { return m_size; }
  void close() { if (m_data) munmap((void*)m_data, m_size); m_data = nullptr; }
  const char* m_data = nullptr; size_t m_size = 0;
}; } }
