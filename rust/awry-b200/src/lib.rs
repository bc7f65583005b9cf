//! `awry_b200::FmIndex`: the reference's query-side API (awry 0.3.1, src/fm_index.rs) on top of
//! the CUDA library.  Signatures are the reference's:
//!   load(&Path) -> Result<FmIndex, io::Error>                     fm_index_file.rs:132
//!   new(&FmBuildArgs) -> Result<FmIndex, _>  (GPU construction)    fm_index.rs:142
//!   count_string(&self, &str) -> u64                               fm_index.rs:499
//!   locate_string(&self, &str) -> Vec<LocalizedSequencePosition>   fm_index.rs:516
//!   parallel_count(&self, impl ParallelIterator<Item=&str>) -> Vec<u64>                       fm_index.rs:455
//!   parallel_locate(&self, impl ParallelIterator<Item=&str>) -> Vec<Vec<LocalizedSequencePosition>>  fm_index.rs:479
//!   update_range_with_symbol / backstep / initial_search_range     fm_index.rs:383, :559-593
//! NOTE: this crate could not be compiled in the environment this repository was developed in
//! (no cargo/rustc); the same C ABI is exercised from C++ and Python tests instead.
use awry_b200_sys as sys;
use rayon::iter::ParallelIterator;
use std::ffi::{CStr, CString};
use std::io;
use std::path::Path;

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum SymbolAlphabet {
    Nucleotide,
    Amino,
}

#[derive(Clone, Debug, PartialEq, Eq, PartialOrd, Ord, Hash, Default)]
pub struct SearchRange {
    pub start_ptr: u64,
    pub end_ptr: u64,
}
impl SearchRange {
    pub fn zero() -> Self { SearchRange { start_ptr: 1, end_ptr: 0 } }
    pub fn is_empty(&self) -> bool { self.start_ptr > self.end_ptr }
    pub fn len(&self) -> u64 { if self.is_empty() { 0 } else { self.end_ptr - self.start_ptr + 1 } }
    pub fn range_iter(&self) -> core::ops::Range<u64> {
        if self.is_empty() { 0..0 } else { self.start_ptr..(self.end_ptr + 1) }
    }
}

#[derive(Clone, Debug, PartialEq, Eq, PartialOrd, Ord, Hash, Default)]
pub struct LocalizedSequencePosition {
    sequence_idx: usize,
    local_position: usize,
}
impl LocalizedSequencePosition {
    pub fn new(sequence_idx: usize, local_position: usize) -> Self { Self { sequence_idx, local_position } }
    pub fn sequence_idx(&self) -> usize { self.sequence_idx }
    pub fn local_position(&self) -> usize { self.local_position }
}

/// awry::fm_index::FmBuildArgs (fm_index.rs:78-96), field for field
pub struct FmBuildArgs {
    pub input_file_src: std::path::PathBuf,
    pub suffix_array_output_src: Option<std::path::PathBuf>,
    pub suffix_array_compression_ratio: Option<u64>,
    pub lookup_table_kmer_len: Option<u8>,
    pub alphabet: SymbolAlphabet,
    pub max_query_len: Option<usize>,
    pub remove_intermediate_suffix_array_file: bool,
}

pub struct FmIndex {
    handle: *mut sys::awry_index,
    info: sys::awry_info,
}
// the device index is immutable after creation; *_batch calls are thread-safe (include/awry_b200.h)
unsafe impl Send for FmIndex {}
unsafe impl Sync for FmIndex {}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::awry_last_error()).to_string_lossy().into_owned() }
}

impl FmIndex {
    /// FmIndex::load: index replicated on every visible device listed in AWRY_B200_DEVICES ("0,1,..."), default device 0.
    pub fn load(fm_file_src: &Path) -> Result<FmIndex, io::Error> {
        let path = CString::new(fm_file_src.to_string_lossy().as_bytes()).map_err(|e| io::Error::new(io::ErrorKind::InvalidInput, e))?;
        let devices: Vec<i32> = std::env::var("AWRY_B200_DEVICES").ok()
            .map(|s| s.split(',').filter_map(|x| x.trim().parse().ok()).collect()).unwrap_or_default();
        let mut handle = std::ptr::null_mut();
        let rc = unsafe {
            sys::awry_index_load(path.as_ptr(), if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() },
                                 devices.len() as i32, &mut handle)
        };
        if rc != sys::AWRY_OK {
            let kind = if rc == sys::AWRY_ERR_IO { io::ErrorKind::NotFound } else { io::ErrorKind::InvalidData };
            return Err(io::Error::new(kind, last_error()));
        }
        let mut info = unsafe { std::mem::zeroed::<sys::awry_info>() };
        unsafe { sys::awry_index_info(handle, &mut info) };
        Ok(FmIndex { handle, info })
    }

    /// FmIndex::new (fm_index.rs:142-268) on the GPU: FASTA/FASTQ -> searchable index (awry_index_build).
    /// `suffix_array_output_src`, `max_query_len` and `remove_intermediate_suffix_array_file` are libsufr
    /// knobs with no meaning here and are ignored.
    pub fn new(args: &FmBuildArgs) -> Result<FmIndex, io::Error> {
        Self::build(args, None)
    }

    /// `new` followed by `save` (fm_index_file.rs:42) in one call: also writes the `.awry` v1 file.
    pub fn new_and_save(args: &FmBuildArgs, fm_file_src: &Path) -> Result<FmIndex, io::Error> {
        Self::build(args, Some(fm_file_src))
    }

    fn build(args: &FmBuildArgs, save_to: Option<&Path>) -> Result<FmIndex, io::Error> {
        let cstr = |p: &Path| CString::new(p.to_string_lossy().as_bytes()).map_err(|e| io::Error::new(io::ErrorKind::InvalidInput, e));
        let input = cstr(&args.input_file_src)?;
        let output = match save_to { Some(p) => Some(cstr(p)?), None => None };
        let devices: Vec<i32> = std::env::var("AWRY_B200_DEVICES").ok()
            .map(|s| s.split(',').filter_map(|x| x.trim().parse().ok()).collect()).unwrap_or_default();
        let a = sys::awry_build_args {
            input_file_src: input.as_ptr(),
            output_file_src: output.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()),
            alphabet: if args.alphabet == SymbolAlphabet::Nucleotide { 0 } else { 1 },
            lookup_table_kmer_len: args.lookup_table_kmer_len.unwrap_or(0) as u32,
            suffix_array_compression_ratio: args.suffix_array_compression_ratio.unwrap_or(0),
            device: devices.first().copied().unwrap_or(0),
        };
        let mut handle = std::ptr::null_mut();
        let rc = unsafe {
            sys::awry_index_build(&a, if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() },
                                  devices.len() as i32, &mut handle)
        };
        if rc != sys::AWRY_OK {
            let kind = if rc == sys::AWRY_ERR_IO { io::ErrorKind::NotFound } else { io::ErrorKind::InvalidData };
            return Err(io::Error::new(kind, last_error()));
        }
        let mut info = unsafe { std::mem::zeroed::<sys::awry_info>() };
        unsafe { sys::awry_index_info(handle, &mut info) };
        Ok(FmIndex { handle, info })
    }

    /// FmIndex::save (fm_index_file.rs:42-106): writes this index as an `.awry` v1 file (awry_index_save).
    pub fn save(&self, file_output_src: &Path) -> Result<(), io::Error> {
        let path = CString::new(file_output_src.to_string_lossy().as_bytes()).map_err(|e| io::Error::new(io::ErrorKind::InvalidInput, e))?;
        let rc = unsafe { sys::awry_index_save(self.handle, path.as_ptr()) };
        if rc != sys::AWRY_OK { return Err(io::Error::new(io::ErrorKind::Other, last_error())); }
        Ok(())
    }

    /// parallel_count over every record of a FASTQ / FASTA file, parsed on the device (awry_count_reads_file).
    pub fn parallel_count_file(&self, reads: &Path) -> Result<Vec<u64>, io::Error> {
        let path = CString::new(reads.to_string_lossy().as_bytes()).map_err(|e| io::Error::new(io::ErrorKind::InvalidInput, e))?;
        let (mut ptr, mut n) = (std::ptr::null_mut::<u64>(), 0u64);
        let rc = unsafe { sys::awry_count_reads_file(self.handle, path.as_ptr(), &mut ptr, &mut n) };
        if rc != sys::AWRY_OK { return Err(io::Error::new(io::ErrorKind::InvalidData, last_error())); }
        let out = unsafe { std::slice::from_raw_parts(ptr, n as usize) }.to_vec();
        unsafe { sys::awry_buffer_free(ptr as *mut std::ffi::c_void) };
        Ok(out)
    }

    pub fn alphabet(&self) -> SymbolAlphabet { if self.info.alphabet == 0 { SymbolAlphabet::Nucleotide } else { SymbolAlphabet::Amino } }
    pub fn suffix_array_compression_ratio(&self) -> u64 { self.info.sa_ratio }
    pub fn bwt_len(&self) -> u64 { self.info.bwt_len }
    pub fn version_number(&self) -> u64 { self.info.version }
    pub fn prefix_sums(&self) -> Vec<u64> { self.info.prefix_sums[..self.info.n_prefix_sums as usize].to_vec() }

    fn pack<'a>(queries: impl ParallelIterator<Item = &'a str>) -> (Vec<u8>, Vec<u64>) {
        // order-preserving collect, then one contiguous byte buffer + offsets
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

    pub fn parallel_count<'a>(&self, queries: impl ParallelIterator<Item = &'a str>) -> Vec<u64> {
        let (bytes, off) = Self::pack(queries);
        let nq = off.len() - 1;
        let mut counts = vec![0u64; nq];
        let rc = unsafe { sys::awry_count_batch(self.handle, bytes.as_ptr(), off.as_ptr(), nq as u64, counts.as_mut_ptr()) };
        if rc != sys::AWRY_OK { panic!("{}", last_error()); } // the reference panics on the same inputs
        counts
    }

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

    pub fn count_string(&self, query: &str) -> u64 {
        self.parallel_count(rayon::iter::once(query))[0]
    }
    pub fn locate_string(&self, query: &str) -> Vec<LocalizedSequencePosition> {
        self.parallel_locate(rayon::iter::once(query)).pop().unwrap()
    }

    pub fn initial_search_range(&self, ascii_symbol: char) -> SearchRange {
        let mut r = sys::awry_range::default();
        let rc = unsafe { sys::awry_initial_range(self.handle, ascii_symbol as u8, &mut r) };
        if rc != sys::AWRY_OK { panic!("{}", last_error()); }
        SearchRange { start_ptr: r.start_ptr, end_ptr: r.end_ptr }
    }
    pub fn update_range_with_symbol(&self, search_range: SearchRange, ascii_symbol: char) -> SearchRange {
        let mut r = sys::awry_range::default();
        let rc = unsafe {
            sys::awry_update_range(self.handle, sys::awry_range { start_ptr: search_range.start_ptr, end_ptr: search_range.end_ptr },
                                   ascii_symbol as u8, &mut r)
        };
        if rc != sys::AWRY_OK { panic!("{}", last_error()); }
        SearchRange { start_ptr: r.start_ptr, end_ptr: r.end_ptr }
    }
    pub fn backstep(&self, search_pointer: u64) -> u64 {
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
