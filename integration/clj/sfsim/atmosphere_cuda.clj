;; Host-side shim for sfsim: calls libsfsim_atmosphere.so (include/sfsim_atmosphere.h) through coffi / the JDK
;; Foreign Function & Memory API, in the style of src/clj/sfsim/jolt.clj.  Copy to src/clj/sfsim/ in an sfsim
;; checkout (see INTEGRATION.md).  NOT executed in this repository's CI: the build image has no JVM; the
;; Python ctypes layer (sfsim_b200/) binds the same symbols and is what the tests drive.

(ns sfsim.atmosphere-cuda
  "Atmospheric lookup tables computed on an NVIDIA B200 through libsfsim_atmosphere.so"
  (:require
    [coffi.ffi :refer (defcfn) :as ffi]
    [coffi.mem :as mem]
    [sfsim.util :refer (spit-floats)]))


(ffi/load-library "libsfsim_atmosphere.so")


;; ---- structs (include/sfsim_atmosphere.h) -------------------------------------------------------------

(def planet-struct
  [::mem/struct
   [[:centre [::mem/array ::mem/double 3]]
    [:radius ::mem/double]
    [:height ::mem/double]
    [:brightness [::mem/array ::mem/double 3]]]])


(def scatter-struct
  [::mem/struct
   [[:base [::mem/array ::mem/double 3]]
    [:scale ::mem/double]
    [:g ::mem/double]
    [:quotient ::mem/double]]])


(def config-struct
  [::mem/struct
   [[:height-size ::mem/int] [:elevation-size ::mem/int] [:light-elevation-size ::mem/int] [:heading-size ::mem/int]
    [:transmittance-height-size ::mem/int] [:transmittance-elevation-size ::mem/int]
    [:surface-height-size ::mem/int] [:surface-sun-elevation-size ::mem/int]
    [:ray-steps ::mem/int] [:sphere-steps ::mem/int] [:iterations ::mem/int]
    [:padding ::mem/int]
    [:intensity [::mem/array ::mem/double 3]]]])


(defn- planet->c
  "Convert sfsim planet map (atmosphere_lut.clj:24-28) to the C layout"
  [{:sfsim.sphere/keys [centre radius] :as planet}]
  {:centre (vec centre) :radius radius :height (:sfsim.atmosphere/height planet)
   :brightness (vec (:sfsim.atmosphere/brightness planet [0.0 0.0 0.0]))})


(defn- scatter->c
  "Convert sfsim scatter map (atmosphere.clj:35-39) to the C layout; absent keys get the reference's defaults"
  [{:sfsim.atmosphere/keys [scatter-base scatter-scale scatter-g scatter-quotient]}]
  {:base (vec scatter-base) :scale scatter-scale :g (or scatter-g 0.0) :quotient (or scatter-quotient 1.0)})


;; ---- entry points ---------------------------------------------------------------------------------------

(defcfn atmlut-init "Select CUDA device" atmlut_init [::mem/int] ::mem/int)
(defcfn atmlut-destroy "Release the library's device resources" atmlut_destroy [] ::mem/void)
(defcfn atmlut-last-error "Message of the last failed call" atmlut_last_error [] ::mem/c-string)

(defcfn atmlut-device-count "Number of CUDA devices" atmlut_device_count [] ::mem/int)

;; Page-locked host memory for the output tables: FFM arenas hand out pageable memory, which the copy engine cannot
;; write directly (the library would stage the 25 MB through its own pinned buffers: 14 ms instead of 12 ms per build).
(defcfn atmlut-host-alloc "Allocate page-locked host memory" atmlut_host_alloc [::mem/long] ::mem/pointer)
(defcfn atmlut-host-free "Free page-locked host memory" atmlut_host_free [::mem/pointer] ::mem/void)

(defcfn atmlut-generate-
  "generate-atmosphere-luts on the first num-gpus GPUs of the box, driven from this one JVM thread (private)"
  atmlut_generate_multi
  [::mem/pointer ::mem/pointer ::mem/int ::mem/pointer ::mem/int
   ::mem/pointer ::mem/pointer ::mem/pointer ::mem/pointer]
  ::mem/int)


(defn- check
  "Throw if a library call failed (every entry point returns 0 on success)"
  [status]
  (when-not (zero? status)
    (throw (RuntimeException. (str "libsfsim_atmosphere: " (atmlut-last-error))))))


(defn- generate-on
  "One call into the library on num-gpus GPUs; outputs are page-locked segments owned by the library"
  [planet* scatter* n-scatter config* num-gpus sizes]
  (let [outputs (mapv (fn [n] (mem/reinterpret (atmlut-host-alloc (* 4 n)) (* 4 n))) sizes)]
    (try
      (check (apply atmlut-generate- planet* scatter* n-scatter config* num-gpus outputs))
      (mapv (fn [segment n] (float-array (mem/deserialize-from segment [::mem/array ::mem/float n]))) outputs sizes)
      (finally
        (run! atmlut-host-free outputs)))))


(defn generate-tables
  "Compute the four tables on all GPUs of the box (at most 8; one GPU if they cannot access each other's memory);
   returns float arrays in file layout [transmittance surface-radiance ray-scatter mie-strength]"
  [planet scatter {:keys [height-size elevation-size light-elevation-size heading-size
                          transmittance-height-size transmittance-elevation-size
                          surface-height-size surface-sun-elevation-size
                          ray-steps sphere-steps iterations intensity] :as config}]
  (with-open [arena (mem/confined-arena)]
    (let [n-t      (* transmittance-height-size transmittance-elevation-size 3)
          n-e      (* surface-height-size surface-sun-elevation-size 3)
          n-s      (* height-size elevation-size light-elevation-size heading-size 3)
          planet*  (mem/serialize (planet->c planet) planet-struct arena)
          scatter* (mem/serialize (mapv scatter->c scatter) [::mem/array scatter-struct (count scatter)] arena)
          config*  (mem/serialize (assoc (dissoc config :intensity) :padding 0 :intensity (vec intensity))
                                  config-struct arena)
          sizes    [n-t n-e n-s n-s]
          gpus     (min 8 (atmlut-device-count))]
      (try
        (generate-on planet* scatter* (count scatter) config* gpus sizes)
        (catch RuntimeException e
          (if (and (> gpus 1) (re-find #"peer access" (.getMessage e)))
            (generate-on planet* scatter* (count scatter) config* 1 sizes)
            (throw e)))))))


(defn generate-atmosphere-luts
  "Drop-in for sfsim.atmosphere-lut/generate-atmosphere-luts (atmosphere_lut.clj:43-105): same constants, same
   four output files, computed by the CUDA library"
  []
  (let [earth    {:sfsim.sphere/centre [0.0 0.0 0.0] :sfsim.sphere/radius 6378000.0
                  :sfsim.atmosphere/height 35000.0 :sfsim.atmosphere/brightness [0.3 0.3 0.3]}
        mie      #:sfsim.atmosphere{:scatter-base [2e-5 2e-5 2e-5] :scatter-scale 1200.0 :scatter-g 0.76
                                    :scatter-quotient 0.9}
        rayleigh #:sfsim.atmosphere{:scatter-base [5.8e-6 13.5e-6 33.1e-6] :scatter-scale 8000.0}
        config   {:height-size 32 :elevation-size 127 :light-elevation-size 32 :heading-size 8
                  :transmittance-height-size 64 :transmittance-elevation-size 255
                  :surface-height-size 16 :surface-sun-elevation-size 63
                  :ray-steps 100 :sphere-steps 15 :iterations 5 :intensity [1.0 1.0 1.0]}
        [t e s m] (generate-tables earth [mie rayleigh] config)]
    (spit-floats "data/atmosphere/transmittance.scatter"    t)
    (spit-floats "data/atmosphere/surface-radiance.scatter" e)
    (spit-floats "data/atmosphere/ray-scatter.scatter"      s)
    (spit-floats "data/atmosphere/mie-strength.scatter"     m)))
