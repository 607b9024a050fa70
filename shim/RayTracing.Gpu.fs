// RayTracing.Gpu.fs — F# host-side binding of librtfs_b200.so for Smaug123/ray-tracing-fsharp.
//
// STATUS: source only.  This image has no .NET SDK (dotnet / fsc / mono are absent), so this file has
// NOT been compiled or run; it is the binding a maintainer would add, written against the reference's
// types as they are at RayTracing/*.fs.  The same marshalling, field for field, is exercised from
// Python (ray_tracing_fsharp_b200/domain.py `marshal`, native.py) by the test-suite.
//
// Where it goes: into the RayTracing project (RayTracing/RayTracing.fsproj), after Scene.fs in the
// compile order, because it reads the private records `Sphere` (Sphere.fs:302-310) and
// `InfinitePlane` (InfinitePlane.fs:101-106).  It adds `Scene.renderGpu`, with the signature of
// `Scene.render` (Scene.fs:196-204), over a `GpuScene` built by `GpuScene.make : Hittable array -> GpuScene`
// (the counterpart of `Scene.make`, Scene.fs:15-28).
namespace RayTracing

open System
open System.Runtime.CompilerServices
open System.Runtime.InteropServices

#nowarn "9" // StructLayout / NativePtr

/// Mirrors of the structs of include/rtfs_b200.h (field order and types must match exactly).
module internal Native =

    [<Literal>]
    let Lib = "rtfs_b200" // librtfs_b200.so on the loader path

    [<Struct ; StructLayout(LayoutKind.Sequential)>]
    type RtTexture =
        val mutable kind : int
        val mutable colourR : byte
        val mutable colourG : byte
        val mutable colourB : byte
        val mutable pad0 : byte
        val mutable width : int
        val mutable height : int
        val mutable rgb8 : nativeint
        val mutable even : int
        val mutable odd : int
        val mutable gridSize : float
        val mutable mapCentreX : float
        val mutable mapCentreY : float
        val mutable mapCentreZ : float
        val mutable mapRadius : float

    [<Struct ; StructLayout(LayoutKind.Sequential)>]
    type RtHittable =
        val mutable shape : int
        val mutable style : int
        val mutable pX : float
        val mutable pY : float
        val mutable pZ : float
        val mutable nX : float
        val mutable nY : float
        val mutable nZ : float
        val mutable radius : float
        val mutable albedo : float
        val mutable fuzz : float
        val mutable ior : float
        val mutable prob : float
        val mutable texture : int
        val mutable colourR : byte
        val mutable colourG : byte
        val mutable colourB : byte
        val mutable pad0 : byte

    [<Struct ; StructLayout(LayoutKind.Sequential)>]
    type RtCamera =
        val mutable viewOriginX : float
        val mutable viewOriginY : float
        val mutable viewOriginZ : float
        val mutable viewDirX : float
        val mutable viewDirY : float
        val mutable viewDirZ : float
        val mutable xAxisOriginX : float
        val mutable xAxisOriginY : float
        val mutable xAxisOriginZ : float
        val mutable xAxisDirX : float
        val mutable xAxisDirY : float
        val mutable xAxisDirZ : float
        val mutable yAxisDirX : float
        val mutable yAxisDirY : float
        val mutable yAxisDirZ : float
        val mutable viewportWidth : float
        val mutable viewportHeight : float
        val mutable focalLength : float
        val mutable samplesPerPixel : int
        val mutable bounceDepth : int

    [<Struct ; StructLayout(LayoutKind.Sequential)>]
    type RtRenderOpts =
        val mutable seed : uint64
        val mutable adaptive : int
        val mutable mode : int
        val mutable gamma : int
        val mutable flags : int

    [<Struct ; StructLayout(LayoutKind.Sequential)>]
    type RtStats =
        val mutable paths : uint64
        val mutable rays : uint64
        val mutable boxTests : uint64
        val mutable primTests : uint64
        val mutable kernelMs : float
        val mutable totalMs : float
        val mutable pixelsEarlyOut : uint64
        val mutable launches : int
        val mutable pad0 : int
        val mutable mainMs : float
        val mutable mainRays : uint64
        val mutable degeneratePaths : uint64

    [<DllImport(Lib)>]
    extern nativeint rt_last_error ()

    [<DllImport(Lib)>]
    extern int rt_device_count ()

    [<DllImport(Lib)>]
    extern int rt_scene_create (RtHittable[] objects, int nObjects, RtTexture[] textures, int nTextures, int device, nativeint& scene)

    [<DllImport(Lib)>]
    extern void rt_scene_destroy (nativeint scene)

    [<DllImport(Lib)>]
    extern int rt_render (nativeint scene, RtCamera& camera, int maxWidthCoord, int maxHeightCoord, RtRenderOpts& opts, byte[] rgbOut, nativeint sumsOut, RtStats& stats)

    [<DllImport(Lib)>]
    extern int rt_multi_create (RtHittable[] objects, int nObjects, RtTexture[] textures, int nTextures, int[] devices, int nDevices, nativeint& multi)

    [<DllImport(Lib)>]
    extern int rt_multi_render (nativeint multi, RtCamera& camera, int maxWidthCoord, int maxHeightCoord, RtRenderOpts& opts, byte[] rgbOut, nativeint sumsOut, RtStats& stats)

    [<DllImport(Lib)>]
    extern void rt_multi_destroy (nativeint multi)

    // one rank of a one-process-per-GPU job: the library issues the NCCL collectives itself (include/rtfs_b200.h, rt_comm_*)
    [<DllImport(Lib)>]
    extern int rt_comm_unique_id (byte[] idOut)

    [<DllImport(Lib)>]
    extern int rt_comm_create (byte[] id, int rank, int world, int device, nativeint stream, nativeint& comm)

    [<DllImport(Lib)>]
    extern int rt_comm_render (nativeint comm, nativeint scene, RtCamera& camera, int maxWidthCoord, int maxHeightCoord, RtRenderOpts& opts, byte[] rgbOut, nativeint sumsOut, RtStats& stats)

    [<DllImport(Lib)>]
    extern void rt_comm_destroy (nativeint comm)

    let check (rc : int) : unit =
        if rc <> 0 then
            failwithf "librtfs_b200 error %i: %s" rc (Marshal.PtrToStringAnsi (rt_last_error ()))

/// `ParameterisedTexture.toTexture` erases an image / checker texture to a closure (Texture.fs:69-72), which
/// cannot cross the ABI.  `GpuTexture.ofParameterised` builds the same `Texture` and remembers its structure,
/// keyed by the closure object, so that `GpuScene.make` can recover it.
[<RequireQualifiedAccess>]
module GpuTexture =
    let private table =
        ConditionalWeakTable<obj, (float * Point * ParameterisedTexture)> ()

    /// Drop-in for `ParameterisedTexture.toTexture (Sphere.planeMapInverse radius centre) texture`.
    let ofParameterised (radius : float) (centre : Point) (texture : ParameterisedTexture) : Texture =
        let t =
            ParameterisedTexture.toTexture (Sphere.planeMapInverse radius centre) texture

        match t with
        | Texture.Arbitrary f -> table.Add (box f, (radius, centre, texture))
        | Texture.Colour _ -> ()

        t

    /// Samples a `ParameterisedTexture` (typically an `Arbitrary` closure, which cannot run on the GPU) into an
    /// `Image` of `width` x `height` texels.  Texel (x, y) of an Image answers the lookups with
    /// `int ((1 - u) (W - 1)) = x` and `int (v (H - 1)) = y` (Texture.fs:63-67); it receives the texture's value
    /// at the centre of that cell.  An approximation by construction, hence explicit (same as
    /// `ParameterisedTexture.bake` in the Python mirror, domain.py).
    let bake (radius : float) (centre : Point) (width : int) (height : int) (texture : ParameterisedTexture) : ParameterisedTexture =
        let interpret = Sphere.planeMapInverse radius centre

        Array.init
            height
            (fun y ->
                let v = min 1.0 ((float y + 0.5) / float (height - 1))

                Array.init
                    width
                    (fun x ->
                        let u = max 0.0 (1.0 - (float x + 0.5) / float (width - 1))
                        ParameterisedTexture.colourAt interpret texture (Sphere.planeMap radius centre u v)
                    )
            )
        |> ParameterisedTexture.Image

    let internal tryStructure (t : Texture) : (float * Point * ParameterisedTexture) voption =
        match t with
        | Texture.Colour _ -> ValueNone
        | Texture.Arbitrary f ->
            match table.TryGetValue (box f) with
            | true, v -> ValueSome v
            | false, _ ->
                failwith
                    "GPU backend: Texture.Arbitrary is a host closure; build it with GpuTexture.ofParameterised (closures inside it: GpuTexture.bake)"

type GpuScene =
    private
        {
            Handle : nativeint
            IsMulti : bool
            Pins : GCHandle list
        }

    interface IDisposable with
        member this.Dispose () =
            if this.IsMulti then
                Native.rt_multi_destroy this.Handle
            else
                Native.rt_scene_destroy this.Handle

            for p in this.Pins do
                p.Free ()

[<RequireQualifiedAccess>]
module GpuScene =

    let private setPoint (Point (struct (x, y, z))) (h : byref<Native.RtHittable>) =
        h.pX <- x
        h.pY <- y
        h.pZ <- z

    /// Flattens `ParameterisedTexture` into the texture table (children before parents, as domain.py does).
    let rec private addTexture
        (textures : ResizeArray<Native.RtTexture>)
        (pins : ResizeArray<GCHandle>)
        (radius : float)
        (Point (struct (cx, cy, cz)))
        (t : ParameterisedTexture)
        : int
        =
        let mutable e = Native.RtTexture ()
        e.even <- -1
        e.odd <- -1
        e.mapCentreX <- cx
        e.mapCentreY <- cy
        e.mapCentreZ <- cz
        e.mapRadius <- radius

        match t with
        | ParameterisedTexture.Colour p ->
            e.kind <- 0
            e.colourR <- p.Red
            e.colourG <- p.Green
            e.colourB <- p.Blue
        | ParameterisedTexture.Image img ->
            // img.[y].[x], rows already flipped by ofImage (Texture.fs:30-48): copy as row-major RGB8
            let h, w = img.Length, img.[0].Length
            let bytes = Array.zeroCreate<byte> (3 * w * h)

            for y in 0 .. h - 1 do
                for x in 0 .. w - 1 do
                    let p = img.[y].[x]
                    bytes.[3 * (y * w + x)] <- p.Red
                    bytes.[3 * (y * w + x) + 1] <- p.Green
                    bytes.[3 * (y * w + x) + 2] <- p.Blue

            let pin = GCHandle.Alloc (bytes, GCHandleType.Pinned)
            pins.Add pin
            e.kind <- 1
            e.width <- w
            e.height <- h
            e.rgb8 <- pin.AddrOfPinnedObject ()
        | ParameterisedTexture.Checkered (even, odd, gridSize) ->
            e.kind <- 2
            e.even <- addTexture textures pins radius (Point (struct (cx, cy, cz))) even
            e.odd <- addTexture textures pins radius (Point (struct (cx, cy, cz))) odd
            e.gridSize <- gridSize
        | ParameterisedTexture.Arbitrary _ ->
            failwith "GPU backend: ParameterisedTexture.Arbitrary is a host closure; bake it to an Image first"

        textures.Add e
        textures.Count - 1

    let private setTexture
        (textures : ResizeArray<Native.RtTexture>)
        (pins : ResizeArray<GCHandle>)
        (t : Texture)
        (h : byref<Native.RtHittable>)
        =
        h.texture <- -1

        match t with
        | Texture.Colour p ->
            h.colourR <- p.Red
            h.colourG <- p.Green
            h.colourB <- p.Blue
        | Texture.Arbitrary _ ->
            match GpuTexture.tryStructure t with
            | ValueSome (radius, centre, structure) -> h.texture <- addTexture textures pins radius centre structure
            | ValueNone -> ()

    let private setColour (p : Pixel) (h : byref<Native.RtHittable>) =
        h.texture <- -1
        h.colourR <- p.Red
        h.colourG <- p.Green
        h.colourB <- p.Blue

    /// The FloatProducer each material carries (Sphere.fs:21-37) is ignored: the device RNG is keyed per
    /// (seed, pixel, sample, bounce).
    let private marshalSphere textures pins (shape : int) (s : Sphere) : Native.RtHittable =
        let mutable h = Native.RtHittable ()
        h.shape <- shape
        setPoint s.Centre &h
        h.radius <- s.Radius

        match s.Style with
        | SphereStyle.LightSource tex ->
            h.style <- 0
            setTexture textures pins tex &h
        | SphereStyle.LightSourceCap colour ->
            h.style <- 1
            setColour colour &h
        | SphereStyle.PureReflection (albedo, tex) ->
            h.style <- 2
            h.albedo <- float albedo
            setTexture textures pins tex &h
        | SphereStyle.FuzzedReflection (albedo, tex, fuzz, _) ->
            h.style <- 3
            h.albedo <- float albedo
            h.fuzz <- float fuzz
            setTexture textures pins tex &h
        | SphereStyle.LambertReflection (albedo, tex, _) ->
            h.style <- 4
            h.albedo <- float albedo
            setTexture textures pins tex &h
        | SphereStyle.Dielectric (albedo, tex, ior, prob, _) ->
            h.style <- 5
            h.albedo <- float albedo
            h.ior <- float ior
            h.prob <- float prob
            setTexture textures pins tex &h
        | SphereStyle.Glass (albedo, tex, ior, _) ->
            h.style <- 6
            h.albedo <- float albedo
            h.ior <- float ior
            setTexture textures pins tex &h

        h

    let private marshalPlane textures pins (p : InfinitePlane) : Native.RtHittable =
        let mutable h = Native.RtHittable ()
        h.shape <- 2
        setPoint p.Point &h
        let (UnitVector (Vector (struct (nx, ny, nz)))) = p.Normal
        h.nX <- nx
        h.nY <- ny
        h.nZ <- nz

        match p.Style with
        | InfinitePlaneStyle.LightSource tex ->
            h.style <- 0
            setTexture textures pins tex &h
        | InfinitePlaneStyle.PureReflection (albedo, colour) ->
            h.style <- 2
            h.albedo <- float albedo
            setColour colour &h
        | InfinitePlaneStyle.FuzzedReflection (albedo, colour, fuzz, _) ->
            h.style <- 3
            h.albedo <- float albedo
            h.fuzz <- float fuzz
            setColour colour &h
        | InfinitePlaneStyle.LambertReflection (albedo, colour, _) ->
            h.style <- 4
            h.albedo <- float albedo
            setColour colour &h

        h

    /// Counterpart of Scene.make (Scene.fs:15-28).  `devices`: CUDA ordinals; more than one splits every frame.
    let makeOn (devices : int[]) (objects : Hittable array) : GpuScene =
        let textures = ResizeArray<Native.RtTexture> ()
        let pins = ResizeArray<GCHandle> ()

        let hittables =
            objects
            |> Array.map (fun o ->
                match o with
                | Hittable.Sphere s -> marshalSphere textures pins 0 s
                | Hittable.UnboundedSphere s -> marshalSphere textures pins 1 s
                | Hittable.InfinitePlane p -> marshalPlane textures pins p
            )

        let texArr =
            if textures.Count = 0 then Array.zeroCreate 1 else textures.ToArray ()

        let mutable handle = 0n

        if devices.Length > 1 then
            Native.rt_multi_create (hittables, hittables.Length, texArr, textures.Count, devices, devices.Length, &handle)
            |> Native.check
        else
            Native.rt_scene_create (hittables, hittables.Length, texArr, textures.Count, devices.[0], &handle)
            |> Native.check

        // the library copies everything it needs during create, so the pins can be released straight away
        for p in pins do
            p.Free ()

        {
            Handle = handle
            IsMulti = devices.Length > 1
            Pins = []
        }

    let make (objects : Hittable array) : GpuScene = makeOn [| 0 |] objects

[<RequireQualifiedAccess>]
module SceneGpu =

    let private marshalCamera (c : Camera) : Native.RtCamera =
        let mutable n = Native.RtCamera ()
        let (Point (struct (ox, oy, oz))) = Ray.origin c.View
        let (UnitVector (Vector (struct (vx, vy, vz)))) = Ray.vector c.View
        let (Point (struct (xox, xoy, xoz))) = Ray.origin c.ViewportXAxis
        let (UnitVector (Vector (struct (xx, xy, xz)))) = Ray.vector c.ViewportXAxis
        let (UnitVector (Vector (struct (yx, yy, yz)))) = Ray.vector c.ViewportYAxis
        n.viewOriginX <- ox
        n.viewOriginY <- oy
        n.viewOriginZ <- oz
        n.viewDirX <- vx
        n.viewDirY <- vy
        n.viewDirZ <- vz
        n.xAxisOriginX <- xox
        n.xAxisOriginY <- xoy
        n.xAxisOriginZ <- xoz
        n.xAxisDirX <- xx
        n.xAxisDirY <- xy
        n.xAxisDirZ <- xz
        n.yAxisDirX <- yx
        n.yAxisDirY <- yy
        n.yAxisDirZ <- yz
        n.viewportWidth <- c.ViewportWidth
        n.viewportHeight <- c.ViewportHeight
        n.focalLength <- c.FocalLength
        n.samplesPerPixel <- c.SamplesPerPixel
        n.bounceDepth <- c.BounceDepth
        n

    /// Same signature and laziness as Scene.render (Scene.fs:196-236): nothing is traced until a row is forced;
    /// the first forced row renders the whole frame on the GPU; every row reports progress once (Scene.fs:232).
    let render
        (progressIncrement : float<progress> -> unit)
        (_print : string -> unit)
        (maxWidthCoord : int)
        (maxHeightCoord : int)
        (camera : Camera)
        (s : GpuScene)
        : float<progress> * Image
        =
        let rowsIter = 2 * maxHeightCoord + 1
        let colsIter = 2 * maxWidthCoord + 1

        let frame =
            lazy
                (let rgb = Array.zeroCreate<byte> (3 * rowsIter * colsIter)
                 let mutable cam = marshalCamera camera
                 let mutable opts = Native.RtRenderOpts ()
                 // FloatProducer (Random ()) in the reference (Scene.fs:205): a fresh seed per frame
                 opts.seed <- uint64 (Random().NextInt64 ())
                 opts.adaptive <- 1
                 let mutable stats = Native.RtStats ()

                 if s.IsMulti then
                     Native.rt_multi_render (s.Handle, &cam, maxWidthCoord, maxHeightCoord, &opts, rgb, 0n, &stats)
                     |> Native.check
                 else
                     Native.rt_render (s.Handle, &cam, maxWidthCoord, maxHeightCoord, &opts, rgb, 0n, &stats)
                     |> Native.check

                 rgb)

        let rows =
            Seq.init
                rowsIter
                (fun row ->
                    async {
                        let rgb = frame.Force ()

                        let result =
                            Array.init
                                colsIter
                                (fun col ->
                                    let i = 3 * (row * colsIter + col)

                                    {
                                        Red = rgb.[i]
                                        Green = rgb.[i + 1]
                                        Blue = rgb.[i + 2]
                                    }
                                )

                        progressIncrement 1.0<progress>
                        return result
                    }
                )

        1.0<progress> * float rowsIter, Image.make rowsIter colsIter rows
