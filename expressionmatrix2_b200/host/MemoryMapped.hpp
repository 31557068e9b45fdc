// Memory-mapped containers of the hot path, BYTE-COMPATIBLE with the reference's files
// (reference src/MemoryMappedVector.hpp:141-198, src/MemoryMappedObject.hpp Header,
//  src/MemoryMappedVectorOfVectors.hpp:28-35): a 256-byte header
//     {headerSize, objectSize, objectCount, pageCount, fileSize, capacity, magic, pad[25]}
// followed by raw T[]; files are whole 4 KiB pages, mapped MAP_SHARED.  Written from scratch for this
// library; only the on-disk format and the public method names follow the reference, so that the
// reference's CellGraph (src/CellGraph.cpp:33-117) and every other reader keep working unchanged.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <limits>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>

namespace ChanZuckerberg {
namespace ExpressionMatrix2 {

// MurmurHash64A (Austin Appleby, public domain) -- MemoryMapped::Vector::hash uses seed 231
// (reference src/MemoryMappedVector.hpp:715-723).
inline uint64_t murmurHash64A(const void* key, uint64_t len, uint64_t seed)
{
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    const int r = 47;
    uint64_t h = seed ^ (len * m);
    const unsigned char* p = static_cast<const unsigned char*>(key);
    const unsigned char* end = p + (len / 8) * 8;
    for (; p != end; p += 8) {
        uint64_t k;
        std::memcpy(&k, p, 8);
        k *= m;
        k ^= k >> r;
        k *= m;
        h ^= k;
        h *= m;
    }
    uint64_t tail = 0;
    const unsigned rem = unsigned(len & 7);
    for (unsigned i = 0; i < rem; i++) tail |= uint64_t(p[i]) << (8 * i);
    if (rem) {
        h ^= tail;
        h *= m;
    }
    h ^= h >> r;
    h *= m;
    h ^= h >> r;
    return h;
}

namespace MemoryMapped {

namespace detail {

constexpr uint64_t kPage = 4096;
constexpr uint64_t kVectorMagic = 0xa3756fd4b5d8bcc1ULL;
constexpr uint64_t kObjectMagic = 0xb7756f4515d8bc94ULL;

struct FileHeader {
    uint64_t headerSize, objectSize, objectCount, pageCount, fileSize, capacity, magicNumber;
    uint64_t padding[25];
};
static_assert(sizeof(FileHeader) == 256, "the reference header is 256 bytes");

inline FileHeader makeHeader(uint64_t objectSize, uint64_t n, uint64_t requestedCapacity, uint64_t magic)
{
    FileHeader h;
    std::memset(&h, 0, sizeof(h));
    h.headerSize = sizeof(FileHeader);
    h.objectSize = objectSize;
    h.objectCount = n;
    const uint64_t bytes = h.headerSize + objectSize * std::max(requestedCapacity, n);
    h.pageCount = (bytes - 1) / kPage + 1;
    h.fileSize = h.pageCount * kPage;
    h.capacity = (h.fileSize - h.headerSize) / objectSize;
    h.magicNumber = magic;
    return h;
}

// One mapped file.  Owns the mapping, not the interpretation.
class Mapping {
public:
    Mapping() = default;
    Mapping(const Mapping&) = delete;
    Mapping& operator=(const Mapping&) = delete;
    ~Mapping()
    {
        if (base_) closeNoThrow();
    }

    void create(const std::string& name, const FileHeader& h)
    {
        if (base_) throw std::runtime_error("MemoryMapped: " + name_ + " is already open");
        const int fd = ::open(name.c_str(), O_CREAT | O_TRUNC | O_RDWR, S_IRUSR | S_IWUSR | S_IRGRP | S_IROTH);
        if (fd == -1) throw std::runtime_error("Error creating " + name + ": " + std::strerror(errno));
        mapFd(fd, name, h.fileSize, true, true);
        *static_cast<FileHeader*>(base_) = h;
    }

    void open(const std::string& name, bool write)
    {
        if (base_) throw std::runtime_error("MemoryMapped: " + name_ + " is already open");
        const int fd = ::open(name.c_str(), write ? O_RDWR : O_RDONLY);
        if (fd == -1)
            throw std::runtime_error("Error accessing " + name + ": error " + std::to_string(errno) + " " +
                                     std::strerror(errno));
        struct stat st;
        if (::fstat(fd, &st) == -1) {
            ::close(fd);
            throw std::runtime_error("Error accessing " + name + ": fstat failed");
        }
        mapFd(fd, name, uint64_t(st.st_size), write, false);
    }

    // Grow (or shrink) the file to the size in `h` and store `h`; existing data are preserved.
    void remap(const FileHeader& h)
    {
        const std::string name = name_;
        sync();
        unmap();
        const int fd = ::open(name.c_str(), O_RDWR);
        if (fd == -1) throw std::runtime_error("Error reopening " + name);
        mapFd(fd, name, h.fileSize, true, true);
        *static_cast<FileHeader*>(base_) = h;
    }

    void sync()
    {
        if (base_ && writable_ && ::msync(base_, size_, MS_SYNC) == -1)
            throw std::runtime_error("Error during msync for " + name_);
    }
    void close()
    {
        if (!base_) throw std::runtime_error("MemoryMapped: close of a file that is not open");
        sync();
        unmap();
    }
    void remove()
    {
        const std::string name = name_;
        close();
        if (::unlink(name.c_str()) == -1) throw std::runtime_error("Error removing " + name);
    }

    FileHeader* header() const { return static_cast<FileHeader*>(base_); }
    char* payload() const { return static_cast<char*>(base_) + sizeof(FileHeader); }
    bool isOpen() const { return base_ != nullptr; }
    bool writable() const { return writable_; }
    uint64_t mappedSize() const { return size_; }
    const std::string& name() const { return name_; }

private:
    void mapFd(int fd, const std::string& name, uint64_t size, bool write, bool truncate)
    {
        if (truncate && ::ftruncate(fd, off_t(size)) == -1) {
            ::close(fd);
            throw std::runtime_error("Error during ftruncate of " + name);
        }
        void* p = ::mmap(nullptr, size, PROT_READ | (write ? PROT_WRITE : 0), MAP_SHARED, fd, 0);
        ::close(fd);   // the mapping keeps the file alive; descriptors are not hoarded
        if (p == MAP_FAILED) throw std::runtime_error("Error during mmap of " + name);
        base_ = p;
        size_ = size;
        writable_ = write;
        name_ = name;
    }
    void unmap()
    {
        ::munmap(base_, size_);
        base_ = nullptr;
        size_ = 0;
        writable_ = false;
        name_.clear();
    }
    void closeNoThrow() noexcept
    {
        if (writable_) ::msync(base_, size_, MS_SYNC);
        ::munmap(base_, size_);
        base_ = nullptr;
    }

    void* base_ = nullptr;
    uint64_t size_ = 0;
    bool writable_ = false;
    std::string name_;
};

}  // namespace detail

// ------------------------------------------------------------------------------------------------
template <class T> class Vector {
public:
    Vector() = default;

    void createNew(const std::string& name, size_t n = 0, size_t requiredCapacity = 0)
    {
        map_.create(name, detail::makeHeader(sizeof(T), n, requiredCapacity, detail::kVectorMagic));
        for (size_t i = 0; i < n; i++) new (begin() + i) T();
    }
    void accessExisting(const std::string& name, bool readWriteAccess)
    {
        map_.open(name, readWriteAccess);
        const detail::FileHeader* h = map_.header();
        const bool ok = map_.mappedSize() >= sizeof(detail::FileHeader) && h->magicNumber == detail::kVectorMagic &&
                        h->fileSize == map_.mappedSize() && h->objectSize == sizeof(T);
        if (!ok) {
            map_.close();
            throw std::runtime_error("Error accessing " + name + ": not a MemoryMapped::Vector of this element type");
        }
    }
    void accessExistingReadOnly(const std::string& name) { accessExisting(name, false); }
    void accessExistingReadWrite(const std::string& name, bool allowReadOnly)
    {
        if (!allowReadOnly) return accessExisting(name, true);
        try {
            accessExisting(name, true);
        } catch (const std::runtime_error&) {
            accessExisting(name, false);
        }
    }

    void syncToDisk() { map_.sync(); }
    void close() { map_.close(); }
    void remove() { map_.remove(); }

    size_t size() const { return map_.isOpen() ? map_.header()->objectCount : 0; }
    bool empty() const { return size() == 0; }
    size_t capacity() const { return map_.isOpen() ? map_.header()->capacity : 0; }
    T* begin() { return reinterpret_cast<T*>(map_.payload()); }
    const T* begin() const { return reinterpret_cast<const T*>(map_.payload()); }
    T* end() { return begin() + size(); }
    const T* end() const { return begin() + size(); }
    T& operator[](size_t i) { return begin()[i]; }
    const T& operator[](size_t i) const { return begin()[i]; }
    T& front() { return *begin(); }
    T& back() { return *(end() - 1); }
    const T& back() const { return *(end() - 1); }

    void push_back(const T& t)
    {
        resize(size() + 1);
        back() = t;
    }
    void resize(size_t newSize)
    {
        requireWritable();
        const size_t oldSize = size();
        if (newSize > capacity()) {
            // same growth policy as the reference: 1.5x the requested size
            map_.remap(detail::makeHeader(sizeof(T), newSize, size_t(1.5 * double(newSize)), detail::kVectorMagic));
        }
        map_.header()->objectCount = newSize;
        for (size_t i = oldSize; i < newSize; i++) new (begin() + i) T();
    }
    void reserve(size_t newCapacity)
    {
        requireWritable();
        if (newCapacity < size()) throw std::runtime_error("MemoryMapped::Vector::reserve below size");
        if (newCapacity == capacity()) return;
        map_.remap(detail::makeHeader(sizeof(T), size(), newCapacity, detail::kVectorMagic));
    }

    bool operator==(const Vector<T>& that) const
    {
        return size() == that.size() && std::equal(begin(), end(), that.begin());
    }
    uint64_t hash() const
    {
        const uint64_t bytes = size() * sizeof(T);
        if (bytes > uint64_t(std::numeric_limits<int>::max()))
            throw std::runtime_error("MemoryMapped::Vector::hash: more than INT_MAX bytes");   // as the reference
        return murmurHash64A(begin(), bytes, 231);
    }

    bool isOpen() const { return map_.isOpen(); }
    bool isOpenWithWriteAccess() const { return map_.isOpen() && map_.writable(); }
    const std::string& fileName() const { return map_.name(); }

private:
    void requireWritable() const
    {
        if (!isOpenWithWriteAccess()) throw std::runtime_error("MemoryMapped::Vector is not open with write access");
    }
    detail::Mapping map_;
};

// ------------------------------------------------------------------------------------------------
template <class T> class Object {
public:
    void createNew(const std::string& name)
    {
        detail::FileHeader h = detail::makeHeader(sizeof(T), 1, 1, detail::kObjectMagic);
        h.capacity = 1;
        map_.create(name, h);
        new (map_.payload()) T();
    }
    void accessExisting(const std::string& name, bool readWriteAccess)
    {
        map_.open(name, readWriteAccess);
        const detail::FileHeader* h = map_.header();
        if (h->magicNumber != detail::kObjectMagic || h->fileSize != map_.mappedSize() || h->objectSize != sizeof(T)) {
            map_.close();
            throw std::runtime_error("Error accessing " + name + ": not a MemoryMapped::Object of this type");
        }
    }
    void accessExistingReadOnly(const std::string& name) { accessExisting(name, false); }
    void accessExistingReadWrite(const std::string& name) { accessExisting(name, true); }
    void syncToDisk() { map_.sync(); }
    void close() { map_.close(); }
    void remove() { map_.remove(); }
    T* operator->() { return reinterpret_cast<T*>(map_.payload()); }
    const T* operator->() const { return reinterpret_cast<const T*>(map_.payload()); }
    bool isOpen() const { return map_.isOpen(); }

private:
    detail::Mapping map_;
};

// ------------------------------------------------------------------------------------------------
// name.toc = Vector<Int> of row starts (size rows+1), name.data = Vector<T>.
template <class T, class Int> class VectorOfVectors {
public:
    void createNew(const std::string& name)
    {
        toc_.createNew(name + ".toc");
        toc_.push_back(Int(0));
        data_.createNew(name + ".data");
    }
    void accessExisting(const std::string& name, bool readWriteAccess)
    {
        toc_.accessExisting(name + ".toc", readWriteAccess);
        data_.accessExisting(name + ".data", readWriteAccess);
    }
    void accessExistingReadOnly(const std::string& name) { accessExisting(name, false); }
    void accessExistingReadWrite(const std::string& name, bool allowReadOnly)
    {
        toc_.accessExistingReadWrite(name + ".toc", allowReadOnly);
        data_.accessExistingReadWrite(name + ".data", allowReadOnly);
    }
    void close()
    {
        toc_.close();
        data_.close();
    }
    void remove()
    {
        toc_.remove();
        data_.remove();
    }
    bool isOpen() const { return toc_.isOpen(); }
    size_t size() const { return toc_.size() - 1; }
    size_t totalSize() const { return data_.size(); }
    size_t size(size_t i) const { return size_t(toc_[i + 1] - toc_[i]); }
    T* begin(Int i) { return data_.begin() + toc_[i]; }
    const T* begin(Int i) const { return data_.begin() + toc_[i]; }
    T* end(Int i) { return data_.begin() + toc_[i + 1]; }
    const T* end(Int i) const { return data_.begin() + toc_[i + 1]; }
    void appendVector() { toc_.push_back(toc_.back()); }
    void append(const T& t)
    {
        ++toc_.back();
        data_.push_back(t);
    }
    template <class It> void appendVector(It b, It e)
    {
        const size_t n = size_t(e - b);
        const size_t old = data_.size();
        data_.resize(old + n);
        std::copy(b, e, data_.begin() + old);
        toc_.push_back(Int(old + n));
    }
    // raw arrays, as the C-ABI consumes them
    const Int* tocBegin() const { return toc_.begin(); }
    const T* dataBegin() const { return data_.begin(); }

private:
    Vector<Int> toc_;
    Vector<T> data_;
};

}  // namespace MemoryMapped
}  // namespace ExpressionMatrix2
}  // namespace ChanZuckerberg
