// Device-resident vectors with the deal.II vector interface the reference's call sites use
// (SURVEY 8b "Vector API the callers use"): LinearAlgebra::distributed::Vector<double>,
// BlockVector<double> (main.cc:67-70) and LinearAlgebra::ReshapedVector (main.cc:196-275).
// All arithmetic happens in CUDA kernels behind the C ABI (include/spirk_b200.h).
#pragma once
#include <spirk_b200.h>

#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace spirk_host
{
  class Error : public std::runtime_error
  {
  public:
    Error(const std::string &what, int status = -1)
      : std::runtime_error(what)
      , status(status)
    {}
    int status;
  };

  inline void check(int status, const char *what)
  {
    if (status != SPIRK_OK)
      throw Error(std::string(what) + " failed (status " + std::to_string(status) + "): " + spirk_last_error(), status);
  }
#define SPIRK_CHECK(call) ::spirk_host::check((call), #call)

  // One per host thread / GPU.  Owns the C-ABI context and a size-keyed free list of device
  // buffers (the analogue of deal.II's GrowingVectorMemory: the reference re-creates temporary
  // vectors in every vmult, e.g. main.cc:1016-1017, 1112-1113, 1560).
  class Device
  {
  public:
    explicit Device(int device_index = 0)
    {
      SPIRK_CHECK(spirk_ctx_create(&ctx_, device_index));
    }
    ~Device()
    {
      for (auto &kv : pool_)
        for (double *p : kv.second)
          spirk_free(ctx_, p);
      spirk_ctx_destroy(ctx_);
    }
    Device(const Device &)            = delete;
    Device &operator=(const Device &) = delete;

    spirk_ctx *ctx() const { return ctx_; }
    void       sync() const { SPIRK_CHECK(spirk_ctx_sync(ctx_)); }

    double *acquire(long long n)
    {
      auto it = pool_.find(n);
      if (it != pool_.end() && !it->second.empty())
        {
          double *p = it->second.back();
          it->second.pop_back();
          return p;
        }
      double *p = nullptr;
      SPIRK_CHECK(spirk_malloc(ctx_, &p, (size_t)n));
      bytes_allocated += n * 8;
      return p;
    }
    void release(double *p, long long n)
    {
      if (p)
        pool_[n].push_back(p);
    }
    long long bytes_allocated = 0;

    // z-slab levels (spatial partition): every block of `n_owned` entries is allocated with `lo` entries of ghost planes
    // below and `hi` above its owned range (SPIRK_SLAB_PAD_LO / _HI planes); registered by the level operators
    struct Padding
    {
      long long lo = 0, hi = 0;
    };
    void register_padding(long long n_owned, long long lo, long long hi) { paddings_[n_owned] = Padding{lo, hi}; }
    Padding padding_for(long long n) const
    {
      auto it = paddings_.find(n);
      return it == paddings_.end() ? Padding{} : it->second;
    }
    // the column (space) communicator of this process: inner products of vectors on z-slab levels are summed over it
    spirk_comm *column_comm = nullptr;
    spirk_comm *column_comm_for(long long n) const { return paddings_.count(n) ? column_comm : nullptr; }

  private:
    spirk_ctx                               *ctx_ = nullptr;
    std::map<long long, std::vector<double *>> pool_;
    std::map<long long, Padding>               paddings_;
  };

  class Vector
  {
  public:
    using value_type = double;

    Vector() = default;
    virtual ~Vector() { clear(); }
    Vector(const Vector &)            = delete;
    Vector(Vector &&o) noexcept { swap(o); }
    Vector &operator=(Vector &&o) noexcept
    {
      if (this != &o)
        {
          clear();
          swap(o);
        }
      return *this;
    }

    void clear()
    {
      views_.clear();
      if (owns_ && raw_)
        dev_->release(raw_, alloc_);
      data_ = raw_ = nullptr, size_ = alloc_ = 0, owns_ = false;
    }

    // allocate n_blocks blocks of block_size entries (zeroed unless omitted); contiguous unless the device has ghost-plane
    // padding registered for this block size (z-slab levels): then block b starts at data() + b * stride()
    void reinit(Device &dev, long long block_size, int n_blocks = 1, bool omit_zeroing_entries = false)
    {
      const long long      n   = block_size * n_blocks;
      const Device::Padding pad = dev.padding_for(block_size);
      const long long      st  = block_size + pad.lo + pad.hi;
      if (!(owns_ && dev_ == &dev && size_ == n && stride_ == st))
        {
          clear();
          dev_   = &dev;
          alloc_ = st * n_blocks;
          raw_   = dev.acquire(alloc_);
          data_  = raw_ + pad.lo;
          size_ = n, owns_ = true;
        }
      n_blocks_ = n_blocks, block_size_ = block_size, stride_ = st;
      column_comm_ = dev.column_comm_for(block_size);
      views_.clear();
      if (!omit_zeroing_entries)
        *this = 0.0;
    }
    void reinit(const Vector &other, bool omit_zeroing_entries = false)
    {
      reinit(*other.dev_, other.block_size_, other.n_blocks_, omit_zeroing_entries);
      reduction_comm_ = other.reduction_comm_;
    }
    // non-owning view (blocks at `stride`, 0 = contiguous)
    void view(Device &dev, double *data, long long block_size, int n_blocks = 1, long long stride = 0)
    {
      clear();
      dev_ = &dev, data_ = data, size_ = block_size * n_blocks, owns_ = false;
      n_blocks_ = n_blocks, block_size_ = block_size, stride_ = stride ? stride : block_size;
      column_comm_ = dev.column_comm_for(block_size);
    }

    // entries from data() that pointwise operations run over: contiguous vectors as they are; padded block vectors
    // including the ghost planes between the blocks (finite values, never read as results)
    long long span() const { return (long long)(n_blocks_ - 1) * stride_ + block_size_; }

    Vector &operator=(const double s)
    {
      SPIRK_CHECK(spirk_vec_set(ctx(), data_, span(), s));
      return *this;
    }
    Vector &operator=(const Vector &V)
    {
      if (this == &V)
        return *this;
      if (size_ != V.size_ || !data_)
        reinit(V, true);
      if (stride_ == V.stride_)
        SPIRK_CHECK(spirk_vec_copy(ctx(), data_, V.data_, span()));
      else
        for (int b = 0; b < n_blocks_; ++b)
          SPIRK_CHECK(spirk_vec_copy(ctx(), data_ + b * stride_, V.data_ + b * V.stride_, block_size_));
      return *this;
    }
    void copy_locally_owned_data_from(const Vector &V) { *this = V; }

    void add(double a, const Vector &V)
    {
      same_layout(V);
      SPIRK_CHECK(spirk_vec_axpy(ctx(), data_, a, V.data_, span()));
    }
    void add(double a, const Vector &V, double b, const Vector &W)
    {
      same_layout(V), same_layout(W);
      SPIRK_CHECK(spirk_vec_add2(ctx(), data_, a, V.data_, b, W.data_, span()));
    }
    void sadd(double s, double a, const Vector &V)
    {
      same_layout(V);
      SPIRK_CHECK(spirk_vec_sadd(ctx(), data_, s, a, V.data_, span()));
    }
    void equ(double a, const Vector &V)
    {
      same_layout(V);
      SPIRK_CHECK(spirk_vec_equ(ctx(), data_, a, V.data_, span()));
    }
    Vector &operator*=(double a)
    {
      SPIRK_CHECK(spirk_vec_scale(ctx(), data_, span(), a));
      return *this;
    }
    Vector &operator+=(const Vector &V)
    {
      add(1.0, V);
      return *this;
    }
    Vector &operator-=(const Vector &V)
    {
      add(-1.0, V);
      return *this;
    }

    // reductions; a ReshapedVector additionally sums over its row communicator (main.cc:237-264)
    double operator*(const Vector &V) const
    {
      double     r = 0;
      CommGuard g(*this);
      same_layout(V);
      if (contiguous())
        SPIRK_CHECK(spirk_vec_dot(ctx(), data_, V.data_, size_, &r));
      else
        SPIRK_CHECK(spirk_vec_dot_strided(ctx(), data_, V.data_, block_size_, n_blocks_, stride_, &r));
      return r;
    }
    double norm_sqr() const { return (*this) * (*this); }
    double l2_norm() const { return std::sqrt(norm_sqr()); }
    double add_and_dot(double a, const Vector &V, const Vector &W)
    {
      double     r = 0;
      CommGuard g(*this);
      same_layout(V), same_layout(W);
      if (contiguous())
        SPIRK_CHECK(spirk_vec_add_and_dot(ctx(), data_, a, V.data_, W.data_, size_, &r));
      else
        SPIRK_CHECK(spirk_vec_add_and_dot_strided(ctx(), data_, a, V.data_, W.data_, block_size_, n_blocks_, stride_, &r));
      return r;
    }
    double mean_value() const
    {
      double r = 0;
      if (contiguous())
        SPIRK_CHECK(spirk_vec_sum(ctx(), data_, size_, &r));
      else
        SPIRK_CHECK(spirk_vec_sum_strided(ctx(), data_, block_size_, n_blocks_, stride_, &r));
      return r / (double)size_;
    }
    bool all_zero() const { return norm_sqr() == 0.0; }

    long long size() const { return size_; }
    long long locally_owned_size() const { return size_; }
    double   *get_values() { return data_; }
    double   *data() { return data_; }
    const double *data() const { return data_; }
    Device   &device() const { return *dev_; }
    spirk_ctx *ctx() const { return dev_->ctx(); }
    bool      empty() const { return data_ == nullptr; }

    // block interface (BlockVector)
    unsigned int n_blocks() const { return n_blocks_; }
    long long    block_size() const { return block_size_; }
    long long    stride() const { return stride_; }              // distance between the first entries of consecutive blocks
    bool         contiguous() const { return n_blocks_ == 1 || stride_ == block_size_; }
    Vector      &block(unsigned int i)
    {
      make_views();
      return *views_[i];
    }
    const Vector &block(unsigned int i) const
    {
      const_cast<Vector *>(this)->make_views();
      return *views_[i];
    }
    void collect_sizes() {}

    // ghost handling is a no-op: one GPU holds whole stage vectors
    void update_ghost_values() const {}
    void zero_out_ghost_values() const {}

    // the communicator inner products are summed over: an explicit one (ReshapedVector: the stage communicator, or stage x
    // space), else the column communicator of a z-slab level, else none
    void set_reduction_comm(spirk_comm *c) { reduction_comm_ = c; }
    spirk_comm *reduction_comm() const { return effective_comm(); }
    spirk_comm *effective_comm() const { return reduction_comm_ ? reduction_comm_ : column_comm_; }

    void copy_to_host(double *host) const
    {
      if (contiguous())
        SPIRK_CHECK(spirk_copy_d2h(ctx(), host, data_, (size_t)size_));
      else
        for (int b = 0; b < n_blocks_; ++b)
          SPIRK_CHECK(spirk_copy_d2h(ctx(), host + b * block_size_, data_ + b * stride_, (size_t)block_size_));
    }
    void copy_from_host(const double *host)
    {
      if (contiguous())
        SPIRK_CHECK(spirk_copy_h2d(ctx(), data_, host, (size_t)size_));
      else
        for (int b = 0; b < n_blocks_; ++b)
          SPIRK_CHECK(spirk_copy_h2d(ctx(), data_ + b * stride_, host + b * block_size_, (size_t)block_size_));
    }
    std::vector<double> to_host() const
    {
      std::vector<double> h((size_t)size_);
      copy_to_host(h.data());
      return h;
    }

    void swap(Vector &o) noexcept
    {
      std::swap(dev_, o.dev_), std::swap(data_, o.data_), std::swap(size_, o.size_), std::swap(owns_, o.owns_);
      std::swap(raw_, o.raw_), std::swap(alloc_, o.alloc_), std::swap(stride_, o.stride_);
      std::swap(n_blocks_, o.n_blocks_), std::swap(block_size_, o.block_size_), std::swap(reduction_comm_, o.reduction_comm_);
      std::swap(column_comm_, o.column_comm_);
      views_.clear(), o.views_.clear();
    }

  protected:
    struct CommGuard
    {
      const Vector &v;
      explicit CommGuard(const Vector &v)
        : v(v)
      {
        if (v.effective_comm())
          spirk_ctx_set_reduction_comm(v.ctx(), v.effective_comm());
      }
      ~CommGuard()
      {
        if (v.effective_comm())
          spirk_ctx_set_reduction_comm(v.ctx(), nullptr);
      }
    };
    void make_views()
    {
      if (views_.size() == (size_t)n_blocks_)
        return;
      views_.clear();
      for (int b = 0; b < n_blocks_; ++b)
        {
          views_.emplace_back(new Vector());
          views_.back()->view(*dev_, data_ + b * stride_, block_size_, 1);
          views_.back()->column_comm_ = column_comm_; // (not the row communicator: a stage block is local to its stage)
        }
    }
    void same_layout(const Vector &V) const
    {
      if (V.size_ != size_ || (n_blocks_ > 1 && V.stride_ != stride_))
        throw Error("vector operation on vectors of different size / block layout");
    }

    Device    *dev_  = nullptr;
    double    *data_ = nullptr, *raw_ = nullptr; // first owned entry; start of the allocation (ghost planes below)
    long long  size_ = 0, alloc_ = 0, stride_ = 0;
    bool       owns_ = false;
    int        n_blocks_ = 1;
    long long  block_size_ = 0;
    spirk_comm *reduction_comm_ = nullptr, *column_comm_ = nullptr;
    std::vector<std::unique_ptr<Vector>> views_;
  };

  using VectorType      = Vector;
  using BlockVectorType = Vector; // a block vector is a Vector with n_blocks() > 1 (blocks contiguous)

  // LinearAlgebra::ReshapedVector (main.cc:196-275): a vector whose inner products are also summed
  // over the stage ("row") communicator so that deal.II-style Krylov solvers see the q-stage vector
  // as one.  Here the extra reduction is an NCCL all-reduce of the device-side scalar.
  class ReshapedVector : public Vector
  {
  public:
    using Vector::reinit;
    using Vector::operator=;
    void reinit(const Vector &V, spirk_comm *row_comm)
    {
      Vector::reinit(V, false);
      reduction_comm_ = row_comm;
    }
    spirk_comm *get_row_mpi_communicator() const { return reduction_comm_; }
  };
} // namespace spirk_host
