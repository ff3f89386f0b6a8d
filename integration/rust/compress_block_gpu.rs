//! UNBUILT sketch: what replaces the body of `compress_block` (src/compression/compress_block.rs:24-67) and, better, the
//! `par_bridge` loop of `compress` (src/compression/compress.rs:125-132) in ohsnyt/bzip2-rust.
use bz2b200_sys::*;
use std::ffi::CStr;

pub struct Gpu(*mut bz2b200_ctx);
unsafe impl Send for Gpu {}            // a context serialises its calls internally

impl Gpu {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        match unsafe { bz2b200_create(device, &mut ctx) } {
            0 => Ok(Gpu(ctx)),
            rc => Err(format!("bz2b200_create: {rc} (no CUDA device?  there is no CPU fallback)")),
        }
    }
    fn err(&self) -> String { unsafe { CStr::from_ptr(bz2b200_last_error(self.0)) }.to_string_lossy().into_owned() }

    /// Drop-in for `compress_block(block, block_crc) -> (Vec<u8>, u8)`: one block per call leaves the GPU mostly idle.
    pub fn compress_block(&self, block: &[u8], block_crc: u32) -> (Vec<u8>, u8) {
        let mut r = self.compress_blocks(&[(block, block_crc)]);
        r.pop().unwrap()
    }

    /// The efficient shape: the blocks the RLE1 iterator (rle1.rs:245-264) yields, ~128 at a time.
    pub fn compress_blocks(&self, blocks: &[(&[u8], u32)]) -> Vec<(Vec<u8>, u8)> {
        let ptrs: Vec<*const u8> = blocks.iter().map(|(b, _)| b.as_ptr()).collect();
        let lens: Vec<u32> = blocks.iter().map(|(b, _)| b.len() as u32).collect();
        let crcs: Vec<u32> = blocks.iter().map(|(_, c)| *c).collect();
        let mut outs: Vec<Vec<u8>> = lens.iter().map(|&n| vec![0u8; n as usize + n as usize / 2 + 4096]).collect();
        let out_ptrs: Vec<*mut u8> = outs.iter_mut().map(|o| o.as_mut_ptr()).collect();
        let caps: Vec<usize> = outs.iter().map(|o| o.len()).collect();
        let mut bits = vec![0u64; blocks.len()];
        let rc = unsafe { bz2b200_compress_blocks(self.0, blocks.len() as i32, ptrs.as_ptr(), lens.as_ptr(), crcs.as_ptr(),
                                                  out_ptrs.as_ptr(), caps.as_ptr(), bits.as_mut_ptr()) };
        assert_eq!(rc, 0, "bz2b200_compress_blocks: {}", self.err());
        outs.into_iter().zip(bits).map(|(mut o, b)| {
            o.truncate(((b + 7) / 8) as usize);
            (o, ((8 - b % 8) % 8) as u8)        // the (bytes, padding) pair BitWriter::add_block expects (bitwriter.rs:77)
        }).collect()
    }

    /// Or bypass both: RLE1, splitting and CRC also run on the GPU, the result is the `.bz2` file's bytes.
    pub fn compress(&self, input: &[u8], level: i32) -> Vec<u8> {
        let mut out = vec![0u8; unsafe { bz2b200_compress_bound(input.len()) }];
        let mut n = 0usize;
        let rc = unsafe { bz2b200_compress_stream(self.0, input.as_ptr(), input.len(), level, out.as_mut_ptr(), out.len(), &mut n) };
        assert_eq!(rc, 0, "bz2b200_compress_stream: {}", self.err());
        out.truncate(n);
        out
    }
}
impl Drop for Gpu { fn drop(&mut self) { unsafe { bz2b200_destroy(self.0) } } }
