//! `awry_b200`: the public API of the reference crate (`awry` 0.3.1) on top of the CUDA library, module for
//! module and signature for signature, so that `awry = { package = "awry-b200", path = ".." }` in a
//! user's Cargo.toml is the whole migration:
//!
//!   awry::fm_index::{FmIndex, FmBuildArgs}             fm_index.rs:41-56, :78-96
//!     FmBuildArgs::new(..)                              fm_index.rs:100
//!     FmIndex::new(&FmBuildArgs) -> Result<Self, anyhow::Error>   (GPU construction)   fm_index.rs:142
//!     FmIndex::load(&Path) -> Result<FmIndex, io::Error>          fm_index_file.rs:132
//!     FmIndex::save(&self, &Path) -> Result<(), io::Error>        fm_index_file.rs:42
//!     count_string(&self, &str) -> u64                            fm_index.rs:499
//!     locate_string(&self, &str) -> Vec<LocalizedSequencePosition>   fm_index.rs:516
//!     parallel_count(&self, impl ParallelIterator<Item=&str>) -> Vec<u64>                          fm_index.rs:455
//!     parallel_locate(&self, impl ParallelIterator<Item=&str>) -> Vec<Vec<LocalizedSequencePosition>>  fm_index.rs:479
//!     initial_search_range(&self, Symbol) / update_range_with_symbol(&self, SearchRange, Symbol) /
//!     backstep(&self, SearchPtr)                                  fm_index.rs:383, :559-593
//!     alphabet / suffix_array_compression_ratio / bwt_len / version_number / prefix_sums   fm_index.rs:302-368
//!   awry::alphabet::{Symbol, SymbolAlphabet}           alphabet.rs:28-31, :74-78, :109-130
//!   awry::search::{SearchRange, SearchPtr}             search.rs:25-81
//!   awry::sequence_index::LocalizedSequencePosition    sequence_index.rs:32-78
//!
//! Everything is also re-exported at the crate root.
//! NOTE: this crate could not be compiled in the environment this repository was developed in
//! (no cargo/rustc); the same C ABI is exercised from C++ and Python tests instead.
use awry_b200_sys as sys;

pub mod alphabet {
    use std::fmt::{self, Display};

    /// alphabet.rs:27-31
    #[derive(Clone, Debug, PartialEq, PartialOrd, Eq, Ord, Hash, Copy)]
    pub enum SymbolAlphabet {
        Nucleotide,
        Amino,
    }

    impl Display for SymbolAlphabet {
        fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
            write!(f, "{}", match self {
                SymbolAlphabet::Nucleotide => "nucleotide",
                SymbolAlphabet::Amino => "amino",
            })
        }
    }

    impl SymbolAlphabet {
        pub(crate) fn alphabet_id(&self) -> u32 {
            match self {
                SymbolAlphabet::Nucleotide => 0,
                SymbolAlphabet::Amino => 1,
            }
        }
        pub(crate) fn cardinality(&self) -> u8 {
            match self {
                SymbolAlphabet::Nucleotide => 6,
                SymbolAlphabet::Amino => 22,
            }
        }
    }

    /// alphabet.rs:74-78.  Only the two encodings a user can construct (ASCII letter, alphabet index); the
    /// bit-vector encoding is crate-private in the reference and lives on the device here.
    #[derive(Debug, PartialEq, Eq)]
    pub struct Symbol {
        alphabet: SymbolAlphabet,
        ascii: char,
    }

    impl Display for Symbol {
        fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
            write!(f, "{} code {}", self.alphabet, self.ascii)
        }
    }

    const NUCLEOTIDE_BY_INDEX: [char; 6] = ['$', 'A', 'C', 'G', 'N', 'T'];
    const AMINO_BY_INDEX: [char; 22] = ['$', 'A', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'K', 'L', 'M', 'N', 'P', 'Q', 'R', 'S',
                                        'T', 'V', 'W', 'X', 'Y'];

    impl Symbol {
        /// alphabet.rs:109-114 (case-insensitive; the library maps the letter to its index, alphabet.rs:169-248)
        pub fn new_ascii(alphabet: SymbolAlphabet, ascii: char) -> Symbol {
            Symbol { alphabet, ascii: ascii.to_ascii_uppercase() }
        }
        /// alphabet.rs:124-130; index -> letter as in to_ascii (alphabet.rs:334-412)
        pub fn new_index(alphabet: SymbolAlphabet, index: u8) -> Symbol {
            debug_assert!(index < alphabet.cardinality());
            let ascii = match alphabet {
                SymbolAlphabet::Nucleotide => *NUCLEOTIDE_BY_INDEX.get(index as usize).unwrap_or(&'N'),
                SymbolAlphabet::Amino => *AMINO_BY_INDEX.get(index as usize).unwrap_or(&'X'),
            };
            Symbol { alphabet, ascii }
        }
        pub(crate) fn ascii_byte(&self) -> u8 {
            if self.ascii.is_ascii() { self.ascii as u8 } else { b'?' } // '?' is searched as N / X, like any unknown letter
        }
    }
}

pub mod search {
    use crate::alphabet::Symbol;
    use crate::fm_index::FmIndex;

    /// search.rs:7 (a crate-private alias of u64 there; public signatures read the same)
    pub type SearchPtr = u64;

    /// search.rs:25-28: inclusive [start_ptr, end_ptr], empty iff start_ptr > end_ptr
    #[derive(Clone, Debug, PartialEq, PartialOrd, Eq, Ord, Hash, Default)]
    pub struct SearchRange {
        pub start_ptr: SearchPtr,
        pub end_ptr: SearchPtr,
    }

    impl SearchRange {
        /// search.rs:43-48: [C[c], C[c+1] - 1]
        pub fn new(fm_index: &FmIndex, symbol: Symbol) -> Self {
            fm_index.initial_search_range(symbol)
        }
        pub fn zero() -> Self {
            SearchRange { start_ptr: 1, end_ptr: 0 }
        }
        #[inline]
        pub fn is_empty(&self) -> bool {
            self.start_ptr > self.end_ptr
        }
        #[inline]
        pub fn len(&self) -> SearchPtr {
            if self.is_empty() { 0 } else { self.end_ptr - self.start_ptr + 1 }
        }
        #[inline]
        pub fn range_iter(&self) -> core::ops::Range<SearchPtr> {
            if self.is_empty() { 0..0 } else { self.start_ptr..(self.end_ptr + 1) }
        }
    }
}

pub mod sequence_index {
    /// sequence_index.rs:32-36
    #[derive(Clone, Debug, PartialEq, Eq, PartialOrd, Ord, Hash, Default)]
    pub struct LocalizedSequencePosition {
        sequence_idx: usize,
        local_position: usize,
    }

    impl LocalizedSequencePosition {
        pub fn new(sequence_idx: usize, local_position: usize) -> Self {
            Self { sequence_idx, local_position }
        }
        pub fn sequence_idx(&self) -> usize {
            self.sequence_idx
        }
        pub fn local_position(&self) -> usize {
            self.local_position
        }
    }
}

pub mod fm_index {
    use super::sys;
    use crate::alphabet::{Symbol, SymbolAlphabet};
    use crate::search::{SearchPtr, SearchRange};
    use crate::sequence_index::LocalizedSequencePosition;
    use rayon::iter::ParallelIterator;
    use std::ffi::{CStr, CString};
    use std::io;
    use std::path::{Path, PathBuf};

    /// fm_index.rs:78-96, field for field.  `suffix_array_output_src`, `max_query_len` and
    /// `remove_intermediate_suffix_array_file` steer libsufr's on-disk suffix sort in the reference; the GPU
    /// construction has no intermediate file and accepts queries of any length, so they are ignored.
    #[derive(Debug)]
    pub struct FmBuildArgs {
        pub input_file_src: PathBuf,
        pub suffix_array_output_src: Option<PathBuf>,
        pub suffix_array_compression_ratio: Option<u64>,
        pub lookup_table_kmer_len: Option<u8>,
        pub alphabet: SymbolAlphabet,
        pub max_query_len: Option<usize>,
        pub remove_intermediate_suffix_array_file: bool,
    }

    impl FmBuildArgs {
        /// fm_index.rs:100-118
        pub fn new(
            input_file_src: PathBuf,
            suffix_array_output_src: Option<PathBuf>,
            suffix_array_compression_ratio: Option<u64>,
            lookup_table_kmer_len: Option<u8>,
            alphabet: SymbolAlphabet,
            max_query_len: Option<usize>,
            remove_intermediate_suffix_array_file: bool,
        ) -> Self {
            FmBuildArgs {
                input_file_src,
                suffix_array_output_src,
                suffix_array_compression_ratio,
                lookup_table_kmer_len,
                alphabet,
                max_query_len,
                remove_intermediate_suffix_array_file,
            }
        }
    }

    /// fm_index.rs:41-56: here a handle to the device replicas
    pub struct FmIndex {
        handle: *mut sys::awry_index,
        info: sys::awry_info,
        prefix_sums: Vec<u64>,
    }
    // the device index is immutable after creation; *_batch calls are thread-safe (include/awry_b200.h)
    unsafe impl Send for FmIndex {}
    unsafe impl Sync for FmIndex {}

    fn last_error() -> String {
        unsafe { CStr::from_ptr(sys::awry_last_error()).to_string_lossy().into_owned() }
    }

    fn c_path(p: &Path) -> Result<CString, io::Error> {
        CString::new(p.to_string_lossy().as_bytes()).map_err(|e| io::Error::new(io::ErrorKind::InvalidInput, e))
    }

    /// devices to replicate the index on: AWRY_B200_DEVICES ("0,1,..."), default device 0
    fn devices() -> Vec<i32> {
        std::env::var("AWRY_B200_DEVICES").ok()
            .map(|s| s.split(',').filter_map(|x| x.trim().parse().ok()).collect()).unwrap_or_default()
    }

    fn io_error(rc: i32) -> io::Error {
        let kind = if rc == sys::AWRY_ERR_IO { io::ErrorKind::NotFound } else { io::ErrorKind::InvalidData };
        io::Error::new(kind, last_error())
    }

    impl FmIndex {
        fn from_handle(handle: *mut sys::awry_index) -> FmIndex {
            let mut info = unsafe { std::mem::zeroed::<sys::awry_info>() };
            unsafe { sys::awry_index_info(handle, &mut info) };
            let prefix_sums = info.prefix_sums[..info.n_prefix_sums as usize].to_vec();
            FmIndex { handle, info, prefix_sums }
        }

        /// FmIndex::load (fm_index_file.rs:132-160)
        pub fn load(fm_file_src: &Path) -> Result<FmIndex, io::Error> {
            let path = c_path(fm_file_src)?;
            let devs = devices();
            let mut handle = std::ptr::null_mut();
            let rc = unsafe {
                sys::awry_index_load(path.as_ptr(), if devs.is_empty() { std::ptr::null() } else { devs.as_ptr() },
                                     devs.len() as i32, &mut handle)
            };
            if rc != sys::AWRY_OK { return Err(io_error(rc)); }
            Ok(Self::from_handle(handle))
        }

        /// FmIndex::new (fm_index.rs:142-268) on the GPU: FASTA/FASTQ -> searchable index (awry_index_build).
        pub fn new(args: &FmBuildArgs) -> Result<Self, anyhow::Error> {
            Self::build(args, None).map_err(anyhow::Error::from)
        }

        /// `new` followed by `save` in one call (the reference-layout arrays are still on the device then,
        /// so nothing has to be re-derived): also writes the `.awry` v1 file.
        pub fn new_and_save(args: &FmBuildArgs, fm_file_src: &Path) -> Result<Self, anyhow::Error> {
            Self::build(args, Some(fm_file_src)).map_err(anyhow::Error::from)
        }

        fn build(args: &FmBuildArgs, save_to: Option<&Path>) -> Result<FmIndex, io::Error> {
            let input = c_path(&args.input_file_src)?;
            let output = match save_to { Some(p) => Some(c_path(p)?), None => None };
            let devs = devices();
            let a = sys::awry_build_args {
                input_file_src: input.as_ptr(),
                output_file_src: output.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()),
                alphabet: args.alphabet.alphabet_id(),
                lookup_table_kmer_len: args.lookup_table_kmer_len.unwrap_or(0) as u32,
                suffix_array_compression_ratio: args.suffix_array_compression_ratio.unwrap_or(0),
                device: devs.first().copied().unwrap_or(0),
            };
            let mut handle = std::ptr::null_mut();
            let rc = unsafe {
                sys::awry_index_build(&a, if devs.is_empty() { std::ptr::null() } else { devs.as_ptr() },
                                      devs.len() as i32, &mut handle)
            };
            if rc != sys::AWRY_OK { return Err(io_error(rc)); }
            Ok(Self::from_handle(handle))
        }

        /// FmIndex::save (fm_index_file.rs:42-106): writes this index as an `.awry` v1 file (awry_index_save).
        pub fn save(&self, file_output_src: &Path) -> Result<(), io::Error> {
            let path = c_path(file_output_src)?;
            let rc = unsafe { sys::awry_index_save(self.handle, path.as_ptr()) };
            if rc != sys::AWRY_OK { return Err(io::Error::new(io::ErrorKind::Other, last_error())); }
            Ok(())
        }

        /// Not in the reference: parallel_count over every record of a FASTQ / FASTA file (plain or gzip), parsed
        /// on the device (awry_count_reads_file).
        pub fn parallel_count_file(&self, reads: &Path) -> Result<Vec<u64>, io::Error> {
            let path = c_path(reads)?;
            let (mut ptr, mut n) = (std::ptr::null_mut::<u64>(), 0u64);
            let rc = unsafe { sys::awry_count_reads_file(self.handle, path.as_ptr(), &mut ptr, &mut n) };
            if rc != sys::AWRY_OK { return Err(io_error(rc)); }
            let out = unsafe { std::slice::from_raw_parts(ptr, n as usize) }.to_vec();
            unsafe { sys::awry_buffer_free(ptr as *mut std::ffi::c_void) };
            Ok(out)
        }

        // getters, fm_index.rs:302-368
        pub fn alphabet(&self) -> SymbolAlphabet {
            if self.info.alphabet == 0 { SymbolAlphabet::Nucleotide } else { SymbolAlphabet::Amino }
        }
        pub fn suffix_array_compression_ratio(&self) -> u64 {
            self.info.sa_ratio
        }
        pub fn bwt_len(&self) -> u64 {
            self.info.bwt_len
        }
        pub fn version_number(&self) -> u64 {
            self.info.version
        }
        pub fn prefix_sums(&self) -> &Vec<u64> {
            &self.prefix_sums
        }

        /// order-preserving collect, then one contiguous byte buffer + offsets (the C ABI's query form)
        fn pack<'a>(queries: impl ParallelIterator<Item = &'a str>) -> (Vec<u8>, Vec<u64>) {
            let qs: Vec<&str> = queries.collect();
            let mut bytes = Vec::with_capacity(qs.iter().map(|q| q.len()).sum::<usize>() + 1);
            let mut off = Vec::with_capacity(qs.len() + 1);
            off.push(0u64);
            for q in qs {
                bytes.extend_from_slice(q.as_bytes());
                off.push(bytes.len() as u64);
            }
            if bytes.is_empty() { bytes.push(0); }
            (bytes, off)
        }

        /// fm_index.rs:455-460
        pub fn parallel_count<'a>(&self, queries: impl ParallelIterator<Item = &'a str>) -> Vec<u64> {
            let (bytes, off) = Self::pack(queries);
            let nq = off.len() - 1;
            let mut counts = vec![0u64; nq];
            let rc = unsafe { sys::awry_count_batch(self.handle, bytes.as_ptr(), off.as_ptr(), nq as u64, counts.as_mut_ptr()) };
            if rc != sys::AWRY_OK { panic!("{}", last_error()); } // the reference panics on the same inputs
            counts
        }

        /// fm_index.rs:479-487; per-query hits in the reference's push order (BWT-row order, fm_index.rs:521)
        pub fn parallel_locate<'a>(&self, queries: impl ParallelIterator<Item = &'a str>) -> Vec<Vec<LocalizedSequencePosition>> {
            let (bytes, off) = Self::pack(queries);
            let nq = off.len() - 1;
            let mut hit_off = vec![0u64; nq + 1];
            let mut hits: *mut sys::awry_hit = std::ptr::null_mut();
            let mut n_hits = 0u64;
            let rc = unsafe {
                sys::awry_locate_batch(self.handle, bytes.as_ptr(), off.as_ptr(), nq as u64, sys::AWRY_LOCATE_BWT_ORDER,
                                       hit_off.as_mut_ptr(), &mut hits, &mut n_hits)
            };
            if rc != sys::AWRY_OK { panic!("{}", last_error()); }
            let all = if n_hits == 0 { &[][..] } else { unsafe { std::slice::from_raw_parts(hits, n_hits as usize) } };
            let out = (0..nq).map(|q| all[hit_off[q] as usize..hit_off[q + 1] as usize].iter()
                .map(|h| LocalizedSequencePosition::new(h.seq_idx as usize, h.local_pos as usize)).collect()).collect();
            unsafe { sys::awry_hits_free(hits) };
            out
        }

        /// fm_index.rs:499-501
        pub fn count_string(&self, query: &str) -> u64 {
            self.parallel_count(rayon::iter::once(query))[0]
        }
        /// fm_index.rs:516-544
        pub fn locate_string(&self, query: &str) -> Vec<LocalizedSequencePosition> {
            self.parallel_locate(rayon::iter::once(query)).pop().unwrap()
        }

        /// fm_index.rs:383
        pub fn initial_search_range(&self, s: Symbol) -> SearchRange {
            let mut r = sys::awry_range::default();
            let rc = unsafe { sys::awry_initial_range(self.handle, s.ascii_byte(), &mut r) };
            if rc != sys::AWRY_OK { panic!("{}", last_error()); }
            SearchRange { start_ptr: r.start_ptr, end_ptr: r.end_ptr }
        }
        /// fm_index.rs:559-582
        pub fn update_range_with_symbol(&self, search_range: SearchRange, query_symbol: Symbol) -> SearchRange {
            let mut r = sys::awry_range::default();
            let rc = unsafe {
                sys::awry_update_range(self.handle, sys::awry_range { start_ptr: search_range.start_ptr, end_ptr: search_range.end_ptr },
                                       query_symbol.ascii_byte(), &mut r)
            };
            if rc != sys::AWRY_OK { panic!("{}", last_error()); }
            SearchRange { start_ptr: r.start_ptr, end_ptr: r.end_ptr }
        }
        /// fm_index.rs:585-593
        pub fn backstep(&self, search_pointer: SearchPtr) -> SearchPtr {
            let mut out = 0u64;
            let rc = unsafe { sys::awry_backstep(self.handle, search_pointer, &mut out) };
            if rc != sys::AWRY_OK { panic!("{}", last_error()); }
            out
        }
    }

    impl Drop for FmIndex {
        fn drop(&mut self) {
            unsafe { sys::awry_index_free(self.handle) }
        }
    }
}

pub use alphabet::{Symbol, SymbolAlphabet};
pub use fm_index::{FmBuildArgs, FmIndex};
pub use search::{SearchPtr, SearchRange};
pub use sequence_index::LocalizedSequencePosition;
