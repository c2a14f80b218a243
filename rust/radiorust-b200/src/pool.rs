//! Pinned-memory chunk pool for the chain edges: the page-locked stand-in for `radiorust::bufferpool`
//! (src/bufferpool.rs:44-222), with the same three types and the same life cycle:
//!
//! | `radiorust::bufferpool` | here | backed by |
//! |---|---|---|
//! | `ChunkBufPool<T>` (`:187-222`) | [`PinnedChunkBufPool<T>`] | `rr_pool_create/get` |
//! | `ChunkBuf<T>` (`:125-165`), `finalize` (`:141`) | [`PinnedChunkBuf<T>`] | one pinned buffer |
//! | `Chunk<T>` (`:44-113`): `Arc` + range + recycler | [`PinnedChunk<T>`] | `rr_pool_put` when the last clone drops |
//!
//! radiorust's `Chunk<T>` is a sealed `Arc<Vec<T>>` (private fields), so foreign memory cannot be handed out as a
//! `Chunk`; the GPU blocks therefore stage through this pool at their two edges (one `memcpy` per direction in the
//! block task) and DMA from/to it asynchronously.  Code that produces samples itself (an SDR reader) can fill a
//! [`PinnedChunkBuf`] directly and hand it to [`crate::chain::Chain::push_pinned`] without that copy.
use crate::{check, sys, Context, Error};

use std::marker::PhantomData;
use std::mem::size_of;
use std::ops::{Deref, DerefMut, Range};
use std::sync::Arc;

struct PoolInner {
    raw: *mut sys::rr_pool,
    _ctx: Context,
}
// rr_pool_get / rr_pool_put lock internally (the reference's recycler is an mpsc channel: any thread may send)
unsafe impl Send for PoolInner {}
unsafe impl Sync for PoolInner {}
impl Drop for PoolInner {
    fn drop(&mut self) {
        unsafe {
            sys::rr_pool_destroy(self.raw);
        }
    }
}

/// One pinned buffer on loan from the pool; goes back through the recycler when dropped
struct Loan {
    pool: Arc<PoolInner>,
    ptr: *mut u8,
    capacity_bytes: usize,
}
unsafe impl Send for Loan {}
unsafe impl Sync for Loan {}
impl Drop for Loan {
    fn drop(&mut self) {
        // bufferpool.rs:82-90: the last owner sends the buffer back to its pool
        unsafe {
            sys::rr_pool_put(self.pool.raw, self.ptr as *mut _);
        }
    }
}

/// Pool to obtain [`PinnedChunkBuf<T>`]s (`ChunkBufPool<T>`, bufferpool.rs:187-222)
pub struct PinnedChunkBufPool<T> {
    inner: Arc<PoolInner>,
    _t: PhantomData<T>,
}

impl<T: Copy> PinnedChunkBufPool<T> {
    /// Create a new pool on `ctx`'s device
    pub fn new(ctx: &Context) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { sys::rr_pool_create(ctx.raw(), &mut raw) })?;
        Ok(Self { inner: Arc::new(PoolInner { raw, _ctx: ctx.clone() }), _t: PhantomData })
    }
    /// Get an empty buffer that can take `capacity` elements: the oldest recycled buffer if it is large enough,
    /// else a new pinned allocation (`ChunkBufPool::get_with_capacity`, bufferpool.rs:210-222)
    pub fn get_with_capacity(&mut self, capacity: usize) -> Result<PinnedChunkBuf<T>, Error> {
        let (mut p, mut cap) = (std::ptr::null_mut(), 0usize);
        check(unsafe { sys::rr_pool_get(self.inner.raw, capacity * size_of::<T>(), &mut p, &mut cap) })?;
        Ok(PinnedChunkBuf {
            loan: Loan { pool: self.inner.clone(), ptr: p as *mut u8, capacity_bytes: cap },
            len: 0,
            _t: PhantomData,
        })
    }
    /// (buffers allocated, buffers reused, idle buffers, buffers on loan)
    pub fn stats(&self) -> (u64, u64, u64, u64) {
        let (mut a, mut r, mut i, mut l) = (0u64, 0u64, 0u64, 0u64);
        unsafe {
            sys::rr_pool_stats(self.inner.raw, &mut a, &mut r, &mut i, &mut l);
        }
        (a, r, i, l)
    }
    /// Free the idle buffers
    pub fn trim(&self) {
        unsafe {
            sys::rr_pool_trim(self.inner.raw);
        }
    }
}

/// Buffer for writing that can be converted into a cheaply cloneable [`PinnedChunk<T>`]
/// (`ChunkBuf<T>`, bufferpool.rs:125-165).  Fixed capacity: pinned memory cannot grow in place.
pub struct PinnedChunkBuf<T> {
    loan: Loan,
    len: usize,
    _t: PhantomData<T>,
}

impl<T: Copy> PinnedChunkBuf<T> {
    /// Elements the buffer can hold
    pub fn capacity(&self) -> usize {
        self.loan.capacity_bytes / size_of::<T>()
    }
    /// Append one element; panics when the capacity is exceeded (a `Vec` would reallocate, pinned memory cannot)
    pub fn push(&mut self, value: T) {
        assert!(self.len < self.capacity(), "capacity of pinned chunk buffer exceeded");
        unsafe {
            (self.loan.ptr as *mut T).add(self.len).write(value);
        }
        self.len += 1;
    }
    /// Append a slice
    pub fn extend_from_slice(&mut self, values: &[T]) {
        assert!(self.len + values.len() <= self.capacity(), "capacity of pinned chunk buffer exceeded");
        unsafe {
            std::ptr::copy_nonoverlapping(values.as_ptr(), (self.loan.ptr as *mut T).add(self.len), values.len());
        }
        self.len += values.len();
    }
    /// Declare the first `len` elements initialised (after the device wrote them)
    ///
    /// # Safety
    /// The caller guarantees that `len` elements have been written (by a completed D2H copy, for example).
    pub unsafe fn set_len(&mut self, len: usize) {
        assert!(len <= self.capacity());
        self.len = len;
    }
    /// Raw pointer for the C ABI
    pub fn as_mut_ptr(&mut self) -> *mut T {
        self.loan.ptr as *mut T
    }
    /// Convert into [`PinnedChunk<T>`] (`ChunkBuf::finalize`, bufferpool.rs:141-143)
    pub fn finalize(self) -> PinnedChunk<T> {
        let len = self.len;
        PinnedChunk { loan: Arc::new(self.loan), range: 0..len, _t: PhantomData }
    }
}

impl<T: Copy> Deref for PinnedChunkBuf<T> {
    type Target = [T];
    fn deref(&self) -> &[T] {
        unsafe { std::slice::from_raw_parts(self.loan.ptr as *const T, self.len) }
    }
}
impl<T: Copy> DerefMut for PinnedChunkBuf<T> {
    fn deref_mut(&mut self) -> &mut [T] {
        unsafe { std::slice::from_raw_parts_mut(self.loan.ptr as *mut T, self.len) }
    }
}

/// Cheaply cloneable, read-only chunk in pinned memory (`Chunk<T>`, bufferpool.rs:44-113).  When the last clone or
/// separated part is dropped the buffer returns to its pool.
#[derive(Clone)]
pub struct PinnedChunk<T> {
    loan: Arc<Loan>,
    range: Range<usize>,
    _t: PhantomData<T>,
}

impl<T: Copy> PinnedChunk<T> {
    /// Discard the first `len` elements (bufferpool.rs:60-63)
    pub fn discard_beginning(&mut self, len: usize) {
        assert!(len <= self.range.end - self.range.start, "length exceeded");
        self.range.start += len;
    }
    /// Split off the first `len` elements without copying (bufferpool.rs:70-79)
    pub fn separate_beginning(&mut self, len: usize) -> Self {
        assert!(len <= self.range.end - self.range.start, "length exceeded");
        let new_range = self.range.start..self.range.start + len;
        self.range.start = new_range.end;
        PinnedChunk { loan: self.loan.clone(), range: new_range, _t: PhantomData }
    }
}

impl<T: Copy> Deref for PinnedChunk<T> {
    type Target = [T];
    fn deref(&self) -> &[T] {
        unsafe { std::slice::from_raw_parts((self.loan.ptr as *const T).add(self.range.start), self.range.end - self.range.start) }
    }
}

/// Pins an existing allocation in place for as long as the guard lives (`rr_host_register`): for `Vec`s that a
/// `radiorust::bufferpool::ChunkBufPool` recycles, so that a registration amortises over many chunks.
pub struct PinnedInPlace {
    ctx: Context,
    ptr: *mut u8,
}
unsafe impl Send for PinnedInPlace {}
impl PinnedInPlace {
    /// # Safety
    /// `data` must stay allocated, and must not be reallocated, until the guard is dropped.
    pub unsafe fn new<T>(ctx: &Context, data: &[T]) -> Result<Self, Error> {
        let ptr = data.as_ptr() as *mut u8;
        check(sys::rr_host_register(ctx.raw(), ptr as *mut _, std::mem::size_of_val(data)))?;
        Ok(Self { ctx: ctx.clone(), ptr })
    }
}
impl Drop for PinnedInPlace {
    fn drop(&mut self) {
        unsafe {
            sys::rr_host_unregister(self.ctx.raw(), self.ptr as *mut _);
        }
    }
}
