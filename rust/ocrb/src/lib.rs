//! Drop-in replacements for the `pub fn`s ocr-rs's two drivers call, backed by libocrb.so.
//! Names, argument meaning and error behaviour follow the reference; `tch::Tensor` arguments become slices /
//! `GrayImage`s because the point of the exercise is to run WITHOUT libtorch.  Not compiled in the build image of this
//! repository (it ships no Rust toolchain): the C ABI underneath is exercised by the Python and C++ mirrors, which are
//! line-for-line the same calls.
use anyhow::{anyhow, Result};
use geo::{LineString, MultiPolygon, Polygon};
use image::GrayImage;
use ocrb_sys as sys;
use std::ffi::{CStr, CString};
use std::path::Path;
use std::ptr;

fn check(rc: i32) -> Result<()> {
    if rc == sys::OCRB_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::ocrb_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow!("libocrb error {}: {}", rc, msg))
}

/// Replaces the process-global `DEVICE` (main.rs:26-28): one context per (device, host thread).
pub struct Ctx(*mut sys::ocrb_ctx);
impl Ctx {
    pub fn new(device: i32) -> Result<Self> {
        let mut p = ptr::null_mut();
        check(unsafe { sys::ocrb_ctx_create(device, &mut p) })?;
        Ok(Ctx(p))
    }
    pub fn raw(&self) -> *mut sys::ocrb_ctx {
        self.0
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::ocrb_ctx_destroy(self.0) };
    }
}

/// `PolygonScores` (metrics.rs:32-35)
pub struct PolygonScores {
    pub polygons: Vec<MultiPolygon<u32>>,
    pub scores: Vec<Vec<f64>>,
    /// ocrb_detect_and_read only: classes of the glyph tiles cut from every polygon, `[polygon][tile]`
    pub glyph_classes: Vec<Vec<Vec<i32>>>,
}

unsafe fn take_polygons(h: *mut sys::ocrb_polygons) -> PolygonScores {
    let nb = sys::ocrb_polygons_num_images(h) as usize;
    let io = std::slice::from_raw_parts(sys::ocrb_polygons_image_offsets(h), nb + 1);
    let npoly = io[nb] as usize;
    let po = std::slice::from_raw_parts(sys::ocrb_polygons_point_offsets(h), npoly + 1);
    let (xy, sc): (&[u32], &[f64]) = if npoly > 0 {
        (std::slice::from_raw_parts(sys::ocrb_polygons_xy(h), 2 * po[npoly] as usize),
         std::slice::from_raw_parts(sys::ocrb_polygons_scores(h), npoly))
    } else {
        (&[], &[])
    };
    let k = sys::ocrb_polygons_glyphs_per_polygon(h) as usize;
    let gc: &[i32] = if k > 0 && npoly > 0 { std::slice::from_raw_parts(sys::ocrb_polygons_glyph_classes(h), npoly * k) } else { &[] };
    let mut out = PolygonScores { polygons: Vec::with_capacity(nb), scores: Vec::with_capacity(nb), glyph_classes: Vec::with_capacity(nb) };
    for b in 0..nb {
        let (mut polys, mut scores, mut classes) = (Vec::new(), Vec::new(), Vec::new());
        for p in io[b] as usize..io[b + 1] as usize {
            let pts: Vec<(u32, u32)> = (po[p] as usize..po[p + 1] as usize).map(|i| (xy[2 * i], xy[2 * i + 1])).collect();
            polys.push(Polygon::new(LineString::from(pts), vec![])); // metrics.rs:111-121
            scores.push(sc[p]);
            if k > 0 {
                classes.push(gc[p * k..(p + 1) * k].to_vec());
            }
        }
        out.polygons.push(MultiPolygon::from(polys));
        out.scores.push(scores);
        out.glyph_classes.push(classes);
    }
    sys::ocrb_polygons_free(h);
    out
}

pub mod image_ops {
    use super::*;
    /// image_ops::preprocess_image (image_ops.rs:188-220), file to padded grey image in one library call: the file is
    /// decoded by libocrb (entropy decode on the host, inverse DCT / upsampling / colour conversion on the GPU,
    /// bit-identical to the `image` crate's decoders), resize + luma + pad run on the decoded pixels in HBM.
    pub fn preprocess_image<T: AsRef<Path>>(ctx: &Ctx, file_path: T, target_dim: (u32, u32)) -> Result<(GrayImage, f64, f64)> {
        let bytes = std::fs::read(file_path.as_ref())?;
        let mut batch = preprocess_images(ctx, &[&bytes[..]], target_dim)?;
        Ok(batch.remove(0))
    }

    /// the same for a batch of encoded files (one decode + one fused resize launch for all of them)
    pub fn preprocess_images(ctx: &Ctx, files: &[&[u8]], target_dim: (u32, u32)) -> Result<Vec<(GrayImage, f64, f64)>> {
        let (w, h) = target_dim;
        let n = files.len();
        let ptrs: Vec<*const u8> = files.iter().map(|f| f.as_ptr()).collect();
        let sizes: Vec<usize> = files.iter().map(|f| f.len()).collect();
        let mut out = vec![0u8; n * (w * h) as usize];
        let mut adjust = vec![0f64; 2 * n];
        check(unsafe {
            sys::ocrb_preprocess_files(ctx.raw(), ptrs.as_ptr(), sizes.as_ptr(), n as i32, w as i32, h as i32, out.as_mut_ptr(), adjust.as_mut_ptr())
        })?;
        Ok(out
            .chunks((w * h) as usize)
            .enumerate()
            .map(|(i, px)| (GrayImage::from_vec(w, h, px.to_vec()).unwrap(), adjust[2 * i], adjust[2 * i + 1]))
            .collect())
    }

    /// image_ops::load_image_as_tensor (image_ops.rs:73-85): open(file)?.into_luma() / 255 as f32 [1, w*h]
    pub fn load_image_as_tensor<T: AsRef<Path>>(ctx: &Ctx, file_path: T) -> Result<Vec<f32>> {
        let path = file_path.as_ref();
        if !path.exists() {
            return Err(anyhow!("File {} doesn't exist", path.display()));
        }
        let bytes = std::fs::read(path)?;
        let (mut w, mut h, mut c) = (0i32, 0i32, 0i32);
        check(unsafe { sys::ocrb_image_info(bytes.as_ptr(), bytes.len(), &mut w, &mut h, &mut c) })?;
        let mut luma = vec![0u8; (w * h) as usize];
        let (ptrs, sizes, offs) = ([bytes.as_ptr()], [bytes.len()], [0i64]);
        check(unsafe { sys::ocrb_decode_images(ctx.raw(), ptrs.as_ptr(), sizes.as_ptr(), 1, sys::OCRB_PIXELS_LUMA, offs.as_ptr(), luma.as_mut_ptr()) })?;
        let mut t = vec![0f32; luma.len()];
        check(unsafe { sys::ocrb_load_image_as_tensor(ctx.raw(), luma.as_ptr(), luma.len() as i64, t.as_mut_ptr()) })?;
        Ok(t)
    }
}

pub mod text_detection {
    use super::*;

    /// `resnet18(&vs.root())` + `vs.load(file)` (model.rs:154, text_detection/mod.rs:35-44): the VarStore archive is read
    /// natively by the library.
    pub struct Resnet18(*mut sys::ocrb_det);
    impl Resnet18 {
        pub fn load<T: AsRef<Path>>(ctx: &Ctx, model_file_path: T, bf16: bool) -> Result<Self> {
            let path = CString::new(model_file_path.as_ref().to_string_lossy().as_bytes())?;
            let mut p = ptr::null_mut();
            check(unsafe { sys::ocrb_det_create_from_file(ctx.raw(), path.as_ptr(), if bf16 { sys::OCRB_MODE_BF16 } else { sys::OCRB_MODE_FP32 }, &mut p) })?;
            Ok(Resnet18(p))
        }
        /// `net.forward_t(&images.view((b, 1, h, w)), false)` (text_detection/mod.rs:52-54): u8 grey levels in, f32 map out
        pub fn forward_t(&self, images: &[u8], b: usize, h: usize, w: usize) -> Result<Vec<f32>> {
            let mut prob = vec![0f32; b * h * w];
            check(unsafe { sys::ocrb_det_forward(self.0, images.as_ptr() as *const _, sys::OCRB_U8, b as i32, h as i32, w as i32, prob.as_mut_ptr()) })?;
            Ok(prob)
        }
        pub fn raw(&self) -> *mut sys::ocrb_det {
            self.0
        }
    }
    impl Drop for Resnet18 {
        fn drop(&mut self) {
            unsafe { sys::ocrb_det_destroy(self.0) };
        }
    }

    pub mod metrics {
        use super::super::*;
        /// metrics.rs:37-56
        pub fn get_boxes_and_box_scores(ctx: &Ctx, pred: &[f32], adjust_values: &[f64], b: usize, h: usize, w: usize) -> Result<PolygonScores> {
            let mut out = ptr::null_mut();
            check(unsafe { sys::ocrb_get_boxes_and_box_scores(ctx.raw(), pred.as_ptr(), adjust_values.as_ptr(), b as i32, h as i32, w as i32, ptr::null(), &mut out) })?;
            Ok(unsafe { take_polygons(out) })
        }
        /// metrics.rs:150-184
        pub fn box_score_fast(ctx: &Ctx, bitmap: &[f32], dim_m2: usize, dim_m1: usize, points: &[(i32, i32)]) -> Result<f64> {
            let flat: Vec<i32> = points.iter().flat_map(|p| vec![p.0, p.1]).collect();
            let mut s = 0f64;
            check(unsafe { sys::ocrb_box_score_fast(ctx.raw(), bitmap.as_ptr(), dim_m2 as i32, dim_m1 as i32, flat.as_ptr(), points.len() as i32, &mut s) })?;
            Ok(s)
        }
        /// metrics.rs:251-372
        pub fn evaluate_image(gt: &MultiPolygon<u32>, ignore_flags: &[bool], pred: &MultiPolygon<u32>) -> Result<sys::ocrb_metrics_item> {
            fn csr(mp: &MultiPolygon<u32>) -> (Vec<i64>, Vec<u32>) {
                let (mut off, mut xy) = (vec![0i64], Vec::new());
                for poly in &mp.0 {
                    let ext = poly.exterior();
                    for p in ext.points_iter().take(ext.num_coords() - 1) {
                        xy.push(p.x());
                        xy.push(p.y());
                    }
                    off.push((xy.len() / 2) as i64);
                }
                (off, xy)
            }
            let (go, gxy) = csr(gt);
            let (do_, dxy) = csr(pred);
            let ig: Vec<u8> = ignore_flags.iter().map(|&f| f as u8).collect();
            let mut item = sys::ocrb_metrics_item::default();
            check(unsafe { sys::ocrb_evaluate_image(go.as_ptr(), gxy.as_ptr(), gt.0.len() as i32, ig.as_ptr(), do_.as_ptr(), dxy.as_ptr(), pred.0.len() as i32, &mut item) })?;
            Ok(item)
        }
    }
}

pub mod utils {
    //! utils.rs:7-79 — the class table, top-k decoding and argument parsing (host code in the reference as well).
    use super::*;

    pub const VALUES: &str = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789";
    pub const VALUES_COUNT: usize = VALUES.len();

    /// `utils::topk` (utils.rs:28-43): the `k` largest of the 62 class scores as `(char, value)`, largest first; equal
    /// values keep the lower class first.  The reference panics on a tensor of another shape; here that is an error.
    pub fn topk(scores: &[f64], k: usize) -> Result<Vec<(char, f64)>> {
        if scores.len() != VALUES_COUNT || k > VALUES_COUNT {
            return Err(anyhow!("unexpected tensor shape [{}]", scores.len()));
        }
        let mut order: Vec<usize> = (0..VALUES_COUNT).collect();
        order.sort_by(|&a, &b| scores[b].partial_cmp(&scores[a]).unwrap_or(std::cmp::Ordering::Equal)); // stable
        Ok(order[..k].iter().map(|&i| (VALUES.as_bytes()[i] as char, scores[i])).collect())
    }

    /// `utils::parse_dimensions` ("800x800", utils.rs:72-79)
    pub fn parse_dimensions(dims_str: &str) -> Result<(u32, u32)> {
        let values: Vec<&str> = dims_str.split_terminator('x').collect();
        let bad = || anyhow!("Could not parse dimensions value: {}", dims_str);
        if values.len() != 2 {
            return Err(bad());
        }
        Ok((values[0].parse().map_err(|_| bad())?, values[1].parse().map_err(|_| bad())?))
    }
}

pub mod polygon {
    use super::*;

    pub enum OffsetType {
        Shrink,
        Expand,
    }

    /// polygon.rs:13-42 — host code (no `Ctx`), as in the reference
    pub fn clip_polygon(polygon: &[(i32, i32)], factor: f64, offset_type: OffsetType) -> Result<Option<Vec<(i32, i32)>>> {
        let flat: Vec<i32> = polygon.iter().flat_map(|p| vec![p.0, p.1]).collect();
        let cap = 6 * polygon.len() + 32;
        let mut out = vec![0i32; 2 * cap];
        let mut n = 0i32;
        let shrink = matches!(offset_type, OffsetType::Shrink) as i32;
        check(unsafe { sys::ocrb_clip_polygon(flat.as_ptr(), polygon.len() as i32, factor, shrink, out.as_mut_ptr(), cap as i32, &mut n, ptr::null_mut()) })?;
        Ok(if n == 0 { None } else { Some((0..n as usize).map(|i| (out[2 * i], out[2 * i + 1])).collect()) })
    }

    /// polygon.rs:44-49
    pub fn shrink_polygon(polygon: &[(i32, i32)], factor: f64) -> Result<Option<Vec<(i32, i32)>>> {
        clip_polygon(polygon, factor, OffsetType::Shrink)
    }

    /// polygon.rs:51-56 (`None` = the empty offset the reference `unwrap()`s)
    pub fn expand_polygon(ctx: &Ctx, polygon: &[(i32, i32)], factor: f64) -> Result<Option<Vec<(i32, i32)>>> {
        let flat: Vec<i32> = polygon.iter().flat_map(|p| vec![p.0, p.1]).collect();
        let cap = 6 * polygon.len() + 32;
        let mut out = vec![0i32; 2 * cap];
        let mut n = 0i32;
        check(unsafe { sys::ocrb_expand_polygon(ctx.raw(), flat.as_ptr(), polygon.len() as i32, factor, out.as_mut_ptr(), cap as i32, &mut n) })?;
        Ok(if n == 0 { None } else { Some((0..n as usize).map(|i| (out[2 * i], out[2 * i + 1])).collect()) })
    }
}

/// `run_text_detection` / the batched evaluation loop (text_detection/mod.rs:23, :188-204) plus recognition of the crops
/// of every detected polygon, over all GPUs of the box: ocrb_detect_and_read_sharded.
pub struct Shards(*mut sys::ocrb_shards);
impl Shards {
    pub fn load<T: AsRef<Path>>(devices: &[i32], det_model: T, rec_model: T, bf16: bool) -> Result<Self> {
        let d = CString::new(det_model.as_ref().to_string_lossy().as_bytes())?;
        let r = CString::new(rec_model.as_ref().to_string_lossy().as_bytes())?;
        let mut p = ptr::null_mut();
        check(unsafe { sys::ocrb_shards_create_from_files(devices.as_ptr(), devices.len() as i32, d.as_ptr(), r.as_ptr(), if bf16 { sys::OCRB_MODE_BF16 } else { sys::OCRB_MODE_FP32 }, &mut p) })?;
        Ok(Shards(p))
    }
    pub fn detect_and_read(&self, images: &[u8], adjust: &[f64], b: usize, h: usize, w: usize, glyphs_per_polygon: i32) -> Result<PolygonScores> {
        let mut out = ptr::null_mut();
        check(unsafe { sys::ocrb_detect_and_read_sharded(self.0, images.as_ptr(), adjust.as_ptr(), b as i32, h as i32, w as i32, ptr::null(), glyphs_per_polygon, &mut out) })?;
        Ok(unsafe { take_polygons(out) })
    }
}
impl Drop for Shards {
    fn drop(&mut self) {
        unsafe { sys::ocrb_shards_destroy(self.0) };
    }
}
